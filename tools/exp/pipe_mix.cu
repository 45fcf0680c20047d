// Experiment: does MUFU.RSQ64H / a DMUL+DADD mix slow the FP64 pipe down?  (build + run on the GPU box)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(double *sink, int iters)
{
    double c[8];
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x + i + 1.0;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 12; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 2 && (rep % 3) == 1) c[i] = c[i] * a;          // DMUL
                else if (MODE == 2 && (rep % 6) == 2) c[i] = c[i] + b;     // DADD
                else c[i] = fma(c[i], a, b);
            }
        }
        if (MODE == 1) {   // 7 MUFU per ~96 FP64 instructions, like the leapfrog step
#pragma unroll
            for (int i = 0; i < 7; i++) {
                double y;
                asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c[i]));
                c[i] = c[i] + y * 1e-300;    // +2 FP64 instr each, counted below
            }
        }
    }
    double t = 0;
    for (int i = 0; i < 8; i++) t += c[i];
    if (t == 123.456) sink[0] = t;
}

template <int MODE> void run(const char *name, double fp64_per_iter)
{
    double *sink; cudaMalloc(&sink, 8);
    int per_sm = 0, dev = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<MODE>, 512, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * per_sm, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int p = 0; p < 2; p++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 512>>>(sink, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = fp64_per_iter * iters * (double)blocks * 512;
    printf("%-28s %8.3f ms  FP64 lane-instr/s %.4g  = %.1f %% of 148*64*1.965e9\n", name, ms, inst / (ms * 1e-3),
           100.0 * inst / (ms * 1e-3) / (148.0 * 64 * 1.965e9));
}

int main()
{
    run<0>("DFMA only", 96);
    run<1>("DFMA + 7 MUFU.RSQ64H /96", 96 + 14);
    run<2>("DFMA/DMUL/DADD mix", 96);
    return 0;
}
