// cw_diag.cu -- pointwise potential kernels (the "next" rows N3: gradient / value / circular
// velocity, replacing potential.gradient / .energy / .circular_velocity at cogsworth
// pop.py:886-895,1001-1013 and classify.py:117-155) and two diagnostics: the DFMA peak
// micro-benchmark that provides the FP64 roofline denominator, and an accuracy self-test of the
// kernels' own rsqrt / reciprocal / log.
#include "cw_kernels.cuh"

namespace cw {

template <int KIND>
__global__ void __launch_bounds__(256) pointwise_kernel(const __grid_constant__ PotParams P, int mode,
                                                        int64_t n, const double *__restrict__ q,
                                                        const double *__restrict__ tt, double *__restrict__ out)
{
    __shared__ double2 ltab_mem[LOG_TABLE_N];
    const LogTab ltab = log_table_init(ltab_mem);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = q[i], y = q[n + i], z = q[2 * n + i];
        const double t = tt ? tt[i] : P.t_eval;      // per-point times [Myr] (time-dependent potentials)
        if (mode == 1) {
            out[i] = potential_value(P, x, y, z, t);
            continue;
        }
        double gx, gy, gz;
        gradient<KIND>(P, ltab, t, x, y, z, gx, gy, gz);
        if (mode == 0) {
            out[i] = gx; out[n + i] = gy; out[2 * n + i] = gz;
        } else {
            // PotentialBase.circular_velocity: sqrt(r |dPhi/dr|), dPhi/dr = (q . grad) / r
            const double r = sqrt(fma(z, z, fma(y, y, x * x)));
            const double dphi_dr = fma(z, gz, fma(y, gy, x * gx)) / r;
            out[i] = sqrt(r * fabs(dphi_dr));
        }
    }
}

cudaError_t launch_pointwise(const PotParams &P, int mode, int64_t n, const double *q, const double *tt, double *out,
                             const LaunchCfg &cfg)
{
    if (n <= 0) return cudaSuccess;
    const int block = 256;
    int64_t grid = (n + block - 1) / block;
    const int64_t cap = (int64_t)cfg.sm_count * 8;
    if (grid > cap) grid = cap;
    switch (P.kind) {
    case POT_MW_V1: pointwise_kernel<POT_MW_V1><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_MW_V2: pointwise_kernel<POT_MW_V2><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_GENERAL: pointwise_kernel<POT_GENERAL><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_ROTATED: pointwise_kernel<POT_ROTATED><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_LM10: pointwise_kernel<POT_LM10><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_MW_V2_T: pointwise_kernel<POT_MW_V2_T><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    case POT_GENERIC_T: pointwise_kernel<POT_GENERIC_T><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    default: pointwise_kernel<POT_GENERIC><<<(unsigned)grid, block, 0, cfg.stream>>>(P, mode, n, q, tt, out); break;
    }
    return cudaGetLastError();
}

// DFMA peak: 8 independent FMA chains per thread, 1024 threads x occupancy-many blocks per SM.
// 16 FMAs per inner iteration; FLOPs = 2 * 16 * iters * threads.
__global__ void __launch_bounds__(512) dfma_peak_kernel(double *sink, int iters)
{
    const double a = 1.0000001, b = 1e-9;
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    const double t = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
    if (t == 123.456) sink[0] = t; // never true; keeps the chains alive
}

cudaError_t launch_dfma_peak(double *sink, int iters, const LaunchCfg &cfg, int *blocks, int *threads)
{
    int per_sm = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dfma_peak_kernel, 512, 0);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    *blocks = cfg.sm_count * per_sm;
    *threads = 512;
    dfma_peak_kernel<<<*blocks, 512, 0, cfg.stream>>>(sink, iters);
    return cudaGetLastError();
}

// max relative error of rsqrt_fast / rcp_fast / log_fast against the CUDA math library over a
// log-uniform sample of the arguments orbits produce (1e-8 .. 1e8; log argument 1 .. 1e4)
__global__ void __launch_bounds__(256) math_selftest_kernel(int64_t n, double *out3)
{
    __shared__ double2 ltab_mem[LOG_TABLE_N];
    const LogTab ltab = log_table_init(ltab_mem);
    PotParams P;
    pot_fill_constants(P);
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // golden-ratio low-discrepancy sequence in [0,1)
        double u = (double)i * 0.6180339887498949;
        u -= floor(u);
        const double x = exp((u * 2.0 - 1.0) * 18.420680743952367); // 1e-8 .. 1e8
        const double w = 1.0 + exp(u * 9.210340371976184) - 1.0 + 1e-12 * (double)(i & 1023); // ~1 .. 1e4
        const double r0 = 1.0 / sqrt(x);
        e0 = fmax(e0, fabs(rsqrt_fast(x) - r0) / r0);
        const double r1 = 1.0 / x;
        e1 = fmax(e1, fabs(rcp_fast(x) - r1) / r1);
        const double r2 = log(w);
        if (r2 > 1e-3) e2 = fmax(e2, fabs(log_fast(w) - r2) / r2);
        if (r2 > 1e-3) e3 = fmax(e3, fabs(log_tab(P, w, ltab) - r2) / r2);
    }
    // orders are preserved for non-negative doubles reinterpreted as integers
    atomicMax((unsigned long long *)&out3[0], (unsigned long long)__double_as_longlong(e0));
    atomicMax((unsigned long long *)&out3[1], (unsigned long long)__double_as_longlong(e1));
    atomicMax((unsigned long long *)&out3[2], (unsigned long long)__double_as_longlong(e2));
    atomicMax((unsigned long long *)&out3[3], (unsigned long long)__double_as_longlong(e3));
}

// Scatter-write micro-benchmark: the HBM write pattern of store_all, without the arithmetic.  `streams`
// independent sequential streams (one per orbit coordinate), each `stream_bytes` long and far apart;
// every visit appends one run of `run_bytes` (16-byte chunks by run_bytes/16 adjacent lanes, one
// STG.128 per lane), all streams advance in lock-step.  Gives the ceiling of the access pattern.
template <int HINT>
__global__ void __launch_bounds__(256) scatter_write_kernel(double2 *base, long long streams, long long stream_bytes,
                                                            int run_bytes)
{
    const int g = run_bytes / 16;                  // lanes per run
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / g; // global run-slot id
    const long long nslots = (long long)gridDim.x * blockDim.x / g;
    const int c = threadIdx.x % g;
    const long long visits = stream_bytes / run_bytes;
    const double2 v = make_double2((double)gw, (double)c);
    unsigned long long pol_keep = 0, pol_drop = 0;
    if (HINT) {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    }
    // visit-major order: every stream receives its k-th run before any stream receives its (k+1)-th,
    // exactly like orbits advancing in lock-step (consecutive runs of one stream are a whole sweep apart)
    for (long long k = 0; k < visits; k++)
        for (long long s0 = gw; s0 < streams; s0 += nslots) {
            double2 *p = base + (s0 * stream_bytes) / 16 + k * g + c;
            if (HINT) {
                // first run of a 128-byte line: keep it in L2 until the second run completes the line
                const bool first_half = (((reinterpret_cast<uintptr_t>(p) / run_bytes) & 1) == 0) && run_bytes == 64;
                const unsigned long long pol = first_half ? pol_keep : pol_drop;
                asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol)
                             : "memory");
            } else {
                *p = v;
            }
        }
}

// The same pattern written by ONE thread per stream: a run of 128 bytes (one full line) as four 32-byte stores
// (STG.256, one full sector per lane and instruction) -- no cooperation between lanes, no shared-memory transposition;
// the L2 receives a sector at a time.  run_bytes code 1128.
__global__ void __launch_bounds__(256) scatter_write_perlane_kernel(double *base, long long streams, long long stream_bytes)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long visits = stream_bytes / 128;
    const double v = (double)tid;
    for (long long k = 0; k < visits; k++)
        for (long long s0 = tid; s0 < streams; s0 += nthreads) {
            double *p = base + (s0 * stream_bytes) / 8 + k * 16;
#pragma unroll
            for (int j = 0; j < 4; j++)
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + 4 * j), "d"(v), "d"(v), "d"(v), "d"(v)
                             : "memory");
        }
}

cudaError_t launch_scatter_write(double2 *base, long long streams, long long stream_bytes, int run_bytes,
                                 int blocks, const LaunchCfg &cfg)
{
    if (run_bytes == 1128) {
        scatter_write_perlane_kernel<<<blocks, 256, 0, cfg.stream>>>((double *)base, streams, stream_bytes);
        return cudaGetLastError();
    }
    if (run_bytes < 0) scatter_write_kernel<1><<<blocks, 256, 0, cfg.stream>>>(base, streams, stream_bytes, -run_bytes);
    else scatter_write_kernel<0><<<blocks, 256, 0, cfg.stream>>>(base, streams, stream_bytes, run_bytes);
    return cudaGetLastError();
}

// dependent-issue latency (cycles) of DFMA, MUFU.RSQ64H+DFMA and the whole rsqrt_fast, one warp alone
__global__ void latency_probe_kernel(double *out)
{
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    const int N = 4096;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, 1.0000001, b);
    long long t1 = clock64();
    double c = 2.0 + a * 1e-30;
#pragma unroll 16
    for (int i = 0; i < N; i++) c = rsqrt_fast(c) + 1.5;
    long long t2 = clock64();
    double d = 2.0 + c * 1e-30;
#pragma unroll 16
    for (int i = 0; i < N; i++) d = d * 1.0000001;
    long long t3 = clock64();
    if (threadIdx.x == 0) {
        out[4] = (double)(t1 - t0) / N;
        out[5] = (double)(t2 - t1) / N;
        out[6] = (double)(t3 - t2) / N;
        out[7] = a + c + d;
    }
}

cudaError_t launch_math_selftest(int64_t n, double *out3, const LaunchCfg &cfg)
{
    cudaError_t err = cudaMemsetAsync(out3, 0, 8 * sizeof(double), cfg.stream);
    if (err != cudaSuccess) return err;
    math_selftest_kernel<<<cfg.sm_count * 4, 256, 0, cfg.stream>>>(n, out3);
    latency_probe_kernel<<<1, 32, 0, cfg.stream>>>(out3);
    return cudaGetLastError();
}

} // namespace cw
