// Experiment 2: do non-FP64 instructions (integer ALU / moves) steal FP64 issue bandwidth on sm_100?
#include <cstdio>
#include <cuda_runtime.h>

template <int NINT>   // NINT integer instructions per 8 DFMA
__global__ void __launch_bounds__(256, 4) k(double *sink, int iters, int seed)
{
    double c[8];
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x + i + 1.0;
    unsigned u[8];
    for (int i = 0; i < 8; i++) u[i] = seed + threadIdx.x * 7 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 12; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
#pragma unroll
            for (int i = 0; i < NINT; i++) u[i % 8] = (u[i % 8] * 1664525u + 1013904223u) ^ (u[(i + 1) % 8] >> 3);
        }
    }
    double t = 0; unsigned v = 0;
    for (int i = 0; i < 8; i++) { t += c[i]; v ^= u[i]; }
    if (t == 123.456 || v == 0x12345u) sink[0] = t + v;
}

template <int NINT> void run()
{
    double *sink; cudaMalloc(&sink, 8);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 4, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int p = 0; p < 2; p++) {
        cudaEventRecord(e0); k<NINT><<<blocks, 256>>>(sink, iters, p); cudaEventRecord(e1); cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 96.0 * iters * (double)blocks * 256;
    printf("8 DFMA + %2d int-ops(x~3 instr): %8.3f ms  FP64 pipe %.1f %%\n", NINT, ms,
           100.0 * inst / (ms * 1e-3) / (148.0 * 64 * 1.965e9));
}

int main() { run<0>(); run<1>(); run<2>(); run<4>(); run<8>(); return 0; }
