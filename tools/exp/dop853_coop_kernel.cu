// EXPERIMENT, NOT BUILT (round 2): measured on a B200 and rejected.  Kept as the record of the measurement.
//   workload (bench.py)        one lane per orbit (shipped)    with this kernel taking the head of the queue
//   config2_dop853             756 ms  (213 003 attempts of ONE orbit x 3.1 us = 660 ms)     1040 ms
//   config5                    141 ms                                                         134 ms
//   config2_dop853_dense       --                                                             518 ms
//   config5_dense              92 ms                                                          76.5 ms
// One attempt of a lone orbit costs 4.9 us here against 3.1 us in dop853_kernel: the attempt is a chain of dependent
// FP64 operations (gradient -> stage -> gradient ...), a third of the stage algebra does not shorten it, and the
// shuffles that carry the position to the four lanes and gather the six error terms (13 x 3 + 4 x 6 per attempt) sit ON
// that chain.  The launch cannot end before its most expensive orbit does, so the workload this was built for got
// slower.  It needs cw_dop853_tableau.cuh (the __constant__ tableau of cw_dop853.cu split into a header) to compile.
//
// cw_dop853_coop.cu -- DOPRI853 for the few orbits that decide the makespan: ONE ORBIT PER GROUP OF FOUR LANES.
//
// An orbit is sequential, so a launch can never end before its most expensive orbit does: attempts x (time of one
// 12-stage attempt).  In dop853_kernel (one orbit per lane, cw_dop853.cu) a lane left alone in its warp still pays the
// full issue cost of every instruction -- an FP64 warp instruction holds the pipe for two cycles however many lanes
// are active -- and measured 3.1 us per attempt (6060 cycles; 213 003 attempts of ONE orbit = the 657 ms of BASELINE
// config 2 with DOPRI853).  Here the six components of an orbit are spread over three lanes (position c and velocity
// c on lane c; the fourth lane mirrors lane 0 and writes nothing), so the stage algebra, the error norm and the
// solution update cost a third of the instructions, and eight orbits share the issue slots of one warp.
//
// THE ARITHMETIC IS THE SAME, BIT FOR BIT, as dop853_kernel's: every component goes through the identical FMA chains,
// every lane evaluates the identical gradient<KIND>() of the same position (exchanged by shuffles inside the group),
// and the two reductions over the six components (error norms, stiffness ratio) gather the six terms and add them in
// dop853_kernel's order.  An orbit therefore returns the same result whichever kernel integrates it
// (tests/test_gpu_parity.py::test_dop853_cooperative_lanes_return_the_same_bits), which keeps results independent
// of batch composition.
//
// Which orbits come here: the pilot launch (cw_dop853.cu) estimates every orbit's step attempts, the queue is sorted
// by that estimate, and dop853_split_kernel hands the head of the queue -- orbits whose own sequential time exceeds a
// fraction of the bulk's makespan -- to this kernel, which runs on its own stream beside dop853_kernel from the start.
//
// Reference semantics: as cw_dop853.cu (gala dop853.pyx + Hairer dop853.c as restated in oracle/cw_oracle.c).
#include <climits>
#include <cstdlib>
#include "cw_kernels.cuh"
#include "cw_dop853_tableau.cuh"

namespace cw {

constexpr int CO_BLOCK = 128;      // 32 groups of four lanes per CTA
constexpr int CO_MAX_CTAS = 37;    // a quarter of the SMs: 1184 groups

// grad at the group's position; returns this lane's component (cc = 0, 1, 2)
template <int KIND>
__device__ __forceinline__ double co_gradient(const PotParams &P, LogTab ltab, double t, double yp, int base, int cc)
{
    const unsigned gm = 0xffffffffu;       // the whole warp, always (see the kernel)
    const double x = __shfl_sync(gm, yp, base + 0), y = __shfl_sync(gm, yp, base + 1), z = __shfl_sync(gm, yp, base + 2);
    double gx, gy, gz;
    gradient<KIND>(P, ltab, t, x, y, z, gx, gy, gz);
    return (cc == 0) ? gx : (cc == 1) ? gy : gz;
}

// sum over the six components in dop853_kernel's order (pos 0, 1, 2, vel 0, 1, 2) of v_i^2, starting from zero
__device__ __forceinline__ double co_sumsq6(double vp, double vv, int base)
{
    const unsigned gm = 0xffffffffu;
    double acc = 0.0;
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const double a = __shfl_sync(gm, vp, base + l);
        acc = fma(a, a, acc);
    }
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const double a = __shfl_sync(gm, vv, base + l);
        acc = fma(a, a, acc);
    }
    return acc;
}

struct CoLane {
    double y[2];      // this lane's components: position cc, velocity cc
    double k1[2];     // f(x, y) of them: (v_cc, -grad_cc)
    double x, xend, h, hmax, facold, hlamb, dt0, tgrid, hgrid, t2, smyr, tend, tlast;
    int64_t e, e1, oid;
    int j, L, next, nstep, naccpt, iasti, nonsti;
    bool last, reject, append;
};

__device__ __forceinline__ void co_apply_kicks(const IntegArgs &A, CoLane &s, int cc)
{
    while (s.e < s.e1 && A.ev_step[s.e] == s.j) {
        s.y[1] += A.ev_dv[3 * s.e + cc];
        const double tk = A.ev_tk[s.e];
        const double tn = (s.append && s.j + 1 == s.L) ? s.t2 : tk + s.hgrid;
        s.dt0 = seg_dt_myr(tk, tn, s.smyr);
        s.e++;
    }
    s.k1[0] = s.y[1];
}

template <bool DENSE>
__device__ __forceinline__ void co_begin_interval(const IntegArgs &A, CoLane &s)
{
    double tn;
    if (DENSE) {
        const bool to_kick = s.e < s.e1;
        s.next = to_kick ? A.ev_step[s.e] : s.L;
        tn = to_kick ? A.ev_tk[s.e] : (s.append ? s.t2 : s.tlast);
        s.tend = tn;
    } else {
        tn = (s.append && s.j + 1 == s.L) ? s.t2 : s.tgrid + s.hgrid;
    }
    s.x = s.tgrid * s.smyr;
    s.xend = tn * s.smyr;
    s.h = s.dt0;
    s.hmax = fabs(s.xend - s.x);
    s.facold = 1.0e-4;
    s.hlamb = 0.0;
    s.nstep = 0; s.naccpt = 0; s.iasti = 0; s.nonsti = 0;
    s.last = false; s.reject = false;
}

__device__ __forceinline__ void co_write_final(const IntegArgs &A, const CoLane &s, int status, int cc, bool writer, bool leader,
                                               int attempts)
{
    const int64_t n = A.n;
    if (writer) {
        A.w_final[cc * n + s.oid] = s.y[0];
        A.w_final[(3 + cc) * n + s.oid] = s.y[1];
    }
    if (leader) {
        if (A.stats && attempts > 0) atomicMax(A.stats + 5, (unsigned long long)attempts);
        if (status != CW_ORBIT_OK) {
            A.status[s.oid] = status;
            if (A.stats) atomicAdd(A.stats + 4, 1ull);
        }
    }
}

template <int KIND, bool DENSE>
__global__ void __launch_bounds__(CO_BLOCK) dop853_coop_kernel(const __grid_constant__ PotParams P,
                                                               const __grid_constant__ IntegArgs A)
{
    const unsigned FULL = 0xffffffffu;
    __shared__ double2 ltab_mem[LOG_TABLE_N];
    const LogTab ltab = log_table_init(ltab_mem);
    const int lane = threadIdx.x & 31, base = lane & ~3, c4 = lane & 3;
    const int cc = (c4 == 3) ? 0 : c4;                 // the fourth lane mirrors lane 0
    const bool writer = c4 < 3, leader = c4 == 0;
    const int64_t M = (int64_t)A.stats[7];             // this kernel's share: queue positions [0, M)
    unsigned long long *queue = A.stats + 8;
    CoLane s;
    s.y[0] = (cc == 0) ? 8.0 : 0.0; s.y[1] = 0.0; s.k1[0] = s.k1[1] = 0.0;
    s.x = 0.0; s.xend = 1.0; s.h = 1.0; s.hmax = 1.0; s.facold = 1e-4; s.hlamb = 0.0; s.dt0 = 1.0;
    s.tgrid = 0.0; s.hgrid = 1.0; s.t2 = 0.0; s.smyr = 1.0; s.tend = 0.0; s.tlast = 0.0;
    s.e = s.e1 = 0; s.oid = -1; s.j = 0; s.L = 0; s.next = 0;
    s.nstep = s.naccpt = s.iasti = s.nonsti = 0;
    s.last = s.reject = s.append = false;
    bool active = false, interval_done = false, exhausted = false;
    int attempts = 0;
    unsigned long long n_acc = 0, n_rej = 0, n_fcn = 0, n_int = 0;
    const int64_t nmax = (A.nmax > 0) ? A.nmax : 100000;
    const double atol = A.atol, rtol = A.rtol;

    // EVERY shuffle below is executed by the whole warp with the full mask, and everything between two shuffles is
    // straight-line code whose side effects are predicated per group.  (Shuffles with one mask per group make the
    // hardware run the groups one after another from then on -- measured 5.2 us per attempt, eight times the issue
    // cost -- so the groups of a warp must never diverge across a shuffle.)
    for (;;) {
        // ---- phase A: the group completed an interval (DENSE: a segment).  No shuffles: may diverge.
        if (active && interval_done) {
            interval_done = false;
            if (DENSE) {
                if (leader) n_int += (unsigned long long)(s.next - s.j);
                s.j = s.next;
                s.tgrid = s.tend;
            } else {
                s.j += 1;
                s.tgrid = (s.append && s.j == s.L) ? s.t2 : s.tgrid + s.hgrid;
                if (leader) n_int++;
            }
            co_apply_kicks(A, s, cc);
            if (s.j == s.L) {
                co_write_final(A, s, CW_ORBIT_OK, cc, writer, leader, attempts);
                active = false;
            } else {
                co_begin_interval<DENSE>(A, s);
            }
        }
        // ---- phase B: idle groups take the next orbits of this kernel's share (warp-uniform trips)
        for (;;) {
            const bool need = !active && !exhausted;
            if (!__any_sync(FULL, need)) break;
            long long p = -1;
            if (need && leader) p = (long long)atomicAdd(queue, 1ull);
            p = __shfl_sync(FULL, p, base);
            bool fresh = false;
            if (need) {
                if (p >= M) {
                    exhausted = true;
                } else {
                    const int64_t o = (int64_t)A.order[p];
                    const int64_t n = A.n;
                    s.oid = o;
                    s.y[0] = A.w0[cc * n + o];
                    s.y[1] = A.w0[(3 + cc) * n + o];
                    s.j = 0;
                    s.L = A.n_nodes[o] - 1;
                    attempts = 0;
                    if (A.status[o] != CW_ORBIT_OK || s.L < 0) {
                        co_write_final(A, s, CW_ORBIT_OK, cc, writer, false, 0);     // status already set by the pre-pass
                    } else {
                        s.hgrid = A.dt[o]; s.t2 = A.t2[o]; s.smyr = A.to_myr[o];
                        s.append = (A.flags[o] & CW_GRID_APPEND_T2) != 0;
                        s.tgrid = A.t1[o];
                        if (DENSE) s.tlast = A.t_last[o];
                        s.e = s.e1 = 0;
                        if (A.ev_off) { s.e = A.ev_off[o] - A.ev_base; s.e1 = A.ev_off[o + 1] - A.ev_base; }
                        const double tn = (s.append && s.L == 1) ? s.t2 : s.tgrid + s.hgrid;
                        s.dt0 = seg_dt_myr(s.tgrid, tn, s.smyr);
                        co_apply_kicks(A, s, cc);
                        if (s.L == 0) co_write_final(A, s, CW_ORBIT_OK, cc, writer, false, 0);
                        else fresh = true;
                    }
                }
            }
            // first derivative of the fresh orbits (the whole warp evaluates; the others discard)
            const double g0 = co_gradient<KIND>(P, ltab, s.tgrid * s.smyr, s.y[0], base, cc);
            if (fresh) {
                s.k1[0] = s.y[1];
                s.k1[1] = -g0;
                if (leader) n_fcn++;
                co_begin_interval<DENSE>(A, s);
                active = true;
            }
        }
        if (!__any_sync(FULL, active)) break;

        // ---- phase C: one step attempt of dopcor(), executed by every lane; `go` predicates the side effects
        int fail = 0;
        if (active) {
            if (s.nstep > nmax) fail = CW_ORBIT_NMAX;
            else if (0.1 * fabs(s.h) <= fabs(s.x) * 2.3e-16) fail = CW_ORBIT_STEP_TOO_SMALL;
            if (fail) {
                co_write_final(A, s, fail, cc, writer, leader, attempts);
                active = false;
            }
        }
        const bool go = active;
        const double posneg = (s.xend - s.x > 0.0) ? 1.0 : -1.0;
        if (go) {
            attempts++;
            if (__dmul_rn(__dsub_rn(__dadd_rn(s.x, __dmul_rn(1.01, s.h)), s.xend), posneg) > 0.0) { s.h = __dsub_rn(s.xend, s.x); s.last = true; }
            s.nstep++;
        }
        const double h = s.h;
        const double *y = s.y, *k1 = s.k1;
        double k2[2], k3[2], k4[2], k5[2], k6[2], k7[2], k8[2], k9[2], k10[2], yy1[2];

        // the stage expressions are dop853_kernel's, component for component
#define CW_STAGE(EXPR) _Pragma("unroll") for (int i = 0; i < 2; i++) yy1[i] = fma(h, (EXPR), y[i]);
#define CW_FCN(T, OUT) { OUT[0] = yy1[1]; OUT[1] = -co_gradient<KIND>(P, ltab, (T), yy1[0], base, cc); }
        CW_STAGE(k_A0201 * k1[i])
        CW_FCN(fma(CW_C02, h, s.x), k2)
        CW_STAGE(fma(k_A0302, k2[i], k_A0301 * k1[i]))
        CW_FCN(fma(CW_C03, h, s.x), k3)
        CW_STAGE(fma(k_A0403, k3[i], k_A0401 * k1[i]))
        CW_FCN(fma(CW_C04, h, s.x), k4)
        CW_STAGE(fma(k_A0504, k4[i], fma(k_A0503, k3[i], k_A0501 * k1[i])))
        CW_FCN(fma(CW_C05, h, s.x), k5)
        CW_STAGE(fma(k_A0605, k5[i], fma(k_A0604, k4[i], k_A0601 * k1[i])))
        CW_FCN(fma(CW_C06, h, s.x), k6)
        CW_STAGE(fma(k_A0706, k6[i], fma(k_A0705, k5[i], fma(k_A0704, k4[i], k_A0701 * k1[i]))))
        CW_FCN(fma(CW_C07, h, s.x), k7)
        CW_STAGE(fma(k_A0807, k7[i], fma(k_A0806, k6[i], fma(k_A0805, k5[i], fma(k_A0804, k4[i], k_A0801 * k1[i])))))
        CW_FCN(fma(CW_C08, h, s.x), k8)
        CW_STAGE(fma(k_A0908, k8[i], fma(k_A0907, k7[i], fma(k_A0906, k6[i], fma(k_A0905, k5[i], fma(k_A0904, k4[i], k_A0901 * k1[i]))))))
        CW_FCN(fma(CW_C09, h, s.x), k9)
        CW_STAGE(fma(k_A1009, k9[i], fma(k_A1008, k8[i], fma(k_A1007, k7[i], fma(k_A1006, k6[i], fma(k_A1005, k5[i], fma(k_A1004, k4[i], k_A1001 * k1[i])))))))
        CW_FCN(fma(CW_C10, h, s.x), k10)
        CW_STAGE(fma(k_A1110, k10[i], fma(k_A1109, k9[i], fma(k_A1108, k8[i], fma(k_A1107, k7[i], fma(k_A1106, k6[i], fma(k_A1105, k5[i], fma(k_A1104, k4[i], k_A1101 * k1[i]))))))))
        CW_FCN(fma(CW_C11, h, s.x), k2)
        const double xph = __dadd_rn(s.x, h);
        CW_STAGE(fma(k_A1211, k2[i], fma(k_A1210, k10[i], fma(k_A1209, k9[i], fma(k_A1208, k8[i], fma(k_A1207, k7[i], fma(k_A1206, k6[i], fma(k_A1205, k5[i], fma(k_A1204, k4[i], k_A1201 * k1[i])))))))))
        CW_FCN(xph, k3)
#undef CW_STAGE
        if (go && leader) n_fcn += 11;

        double sq2[2], sq1[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            k4[i] = fma(k_B12, k3[i], fma(k_B11, k2[i], fma(k_B10, k10[i], fma(k_B09, k9[i],
                    fma(k_B08, k8[i], fma(k_B07, k7[i], fma(k_B06, k6[i], k_B01 * k1[i])))))));
            k5[i] = fma(h, k4[i], y[i]);
            const double sk = fma(rtol, fmax(fabs(y[i]), fabs(k5[i])), atol);
            const double isk = rcp_fast(sk);
            double erri = k4[i] - k_BHH1 * k1[i] - k_BHH2 * k9[i] - k_BHH3 * k3[i];
            sq2[i] = erri * isk;
            erri = fma(k_ER12, k3[i], fma(k_ER11, k2[i], fma(k_ER10, k10[i], fma(k_ER09, k9[i],
                   fma(k_ER08, k8[i], fma(k_ER07, k7[i], fma(k_ER06, k6[i], k_ER01 * k1[i])))))));
            sq1[i] = erri * isk;
        }
        const double err2 = co_sumsq6(sq2[0], sq2[1], base);
        double err = co_sumsq6(sq1[0], sq1[1], base);
        double deno = fma(0.01, err2, err);
        if (!(deno > 1.0e-290)) deno = 1.0;
        err = fabs(h) * err * rsqrt_fast(deno * 6.0);
        const bool nonfinite = !(err <= 1.79769313486231571e+308);
        const double e0 = fmax(err, 1.0e-280);
        const double e1 = e0 * rsqrt_fast(e0);
        const double e2 = e1 * rsqrt_fast(e1);
        const double fac11 = e2 * rsqrt_fast(e2);
        const double facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0, inv_safe = 1.0 / 0.9;
        const double fac = fmax(facc2, fmin(facc1, fac11 * inv_safe));
        double hnew = h * rcp_fast(fac);
        // first stage of the next step f(x + h, ynew) and the stiffness sums: needed by accepted steps only, evaluated
        // by every lane so that the warp stays converged across their shuffles
        double kn[2];
        kn[0] = k5[1];
        kn[1] = -co_gradient<KIND>(P, ltab, xph, k5[0], base, cc);
        const double stnum = co_sumsq6(kn[0] - k3[0], kn[1] - k3[1], base);
        const double stden = co_sumsq6(k5[0] - yy1[0], k5[1] - yy1[1], base);
#undef CW_FCN
        if (go) {
            if (nonfinite) {
                co_write_final(A, s, CW_ORBIT_NONFINITE, cc, writer, leader, attempts);
                active = false;
            } else if (err <= 1.0) {
                // ---- step accepted
                s.facold = fmax(err, 1.0e-4);
                s.naccpt++;
                if (leader) { n_acc++; n_fcn++; }
                if (stden > 0.0) s.hlamb = (h > 0.0 && h * h * stnum > 37.21 * stden) ? 7.0 : 0.0;
                bool stiff = false;
                if (s.hlamb > 6.1) {
                    s.nonsti = 0;
                    s.iasti++;
                    if (s.iasti == 15) stiff = true;
                } else {
                    s.nonsti++;
                    if (s.nonsti == 6) s.iasti = 0;
                }
#pragma unroll
                for (int i = 0; i < 2; i++) { s.k1[i] = kn[i]; s.y[i] = k5[i]; }
                s.x = xph;
                if (stiff) {
                    co_write_final(A, s, CW_ORBIT_STIFF, cc, writer, leader, attempts);
                    active = false;
                } else if (s.last) {
                    interval_done = true;
                } else {
                    if (fabs(hnew) > s.hmax) hnew = posneg * s.hmax;
                    if (s.reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
                    s.reject = false;
                    s.h = hnew;
                }
            } else {
                // ---- step rejected
                hnew = h * rcp_fast(fmin(facc1, fac11 * inv_safe));
                s.reject = true;
                if (s.naccpt >= 1 && leader) n_rej++;
                s.last = false;
                s.h = hnew;
            }
        }
    }
    if (A.stats) {
        for (int off = 16; off > 0; off >>= 1) {
            n_acc += __shfl_down_sync(FULL, n_acc, off);
            n_rej += __shfl_down_sync(FULL, n_rej, off);
            n_fcn += __shfl_down_sync(FULL, n_fcn, off);
            n_int += __shfl_down_sync(FULL, n_int, off);
        }
        if (lane == 0) {
            if (n_acc) atomicAdd(A.stats + 0, n_acc);
            if (n_rej) atomicAdd(A.stats + 1, n_rej);
            if (n_fcn) atomicAdd(A.stats + 2, n_fcn);
            if (n_int) atomicAdd(A.stats + 3, n_int);
        }
    }
}

// The head of the cost-sorted queue goes to the cooperative kernel: orbits whose estimated step attempts exceed
// `frac` x the mean load of a lane of dop853_kernel (their own sequential time would then be a sizeable part of, or
// longer than, the bulk's makespan), at most `cap` of them.  Writes the share M to stats[7], zeroes the cooperative
// queue (stats[8]) and starts dop853_kernel's queue at M.  keys: estimated attempts as float bits, sorted descending.
__global__ void __launch_bounds__(1024) dop853_split_kernel(const __grid_constant__ IntegArgs A, const int32_t *keys,
                                                            int lanes_main, int cap, double frac)
{
    __shared__ double part[32];
    __shared__ double total;
    const int64_t n = A.n;
    double sum = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) sum += fmin((double)__int_as_float(keys[i]), 1.0e9);
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        total = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double thr = frac * total / (double)lanes_main;
        // first position whose cost is <= thr (keys descend)
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            if ((double)__int_as_float(keys[mid]) > thr) lo = mid + 1; else hi = mid;
        }
        int64_t M = lo < (int64_t)cap ? lo : (int64_t)cap;
        if (M < 0) M = 0;
        A.stats[7] = (unsigned long long)M;
        A.stats[8] = 0ull;
        *A.queue = (unsigned long long)M;
    }
}

cudaError_t launch_dop853_split(const IntegArgs &A, const int32_t *sorted_keys, int lanes_main, const LaunchCfg &cfg)
{
    double frac = 0.8;
    if (const char *e = getenv("CW_DOP_COOP_FRAC")) frac = atof(e);
    dop853_split_kernel<<<1, 1024, 0, cfg.stream>>>(A, sorted_keys, lanes_main, 2 * CO_MAX_CTAS * (CO_BLOCK / 4), frac);
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_co_kind(const PotParams &P, const IntegArgs &A, bool dense, const LaunchCfg &cfg)
{
    int grid = cfg.sm_count / 4;
    if (grid > CO_MAX_CTAS) grid = CO_MAX_CTAS;
    if (grid < 1) grid = 1;
    if (dense) dop853_coop_kernel<KIND, true><<<grid, CO_BLOCK, 0, cfg.stream>>>(P, A);
    else dop853_coop_kernel<KIND, false><<<grid, CO_BLOCK, 0, cfg.stream>>>(P, A);
    return cudaGetLastError();
}

cudaError_t launch_dop853_coop(const PotParams &P, const IntegArgs &A, bool dense, const LaunchCfg &cfg)
{
    if (A.n <= 0) return cudaSuccess;
    if (!A.order || (dense && !A.t_last)) return cudaErrorInvalidValue;
    switch (P.kind) {
    case POT_MW_V1: return launch_co_kind<POT_MW_V1>(P, A, dense, cfg);
    case POT_MW_V2: return launch_co_kind<POT_MW_V2>(P, A, dense, cfg);
    case POT_GENERAL: return launch_co_kind<POT_GENERAL>(P, A, dense, cfg);
    case POT_MW_V2_T: return launch_co_kind<POT_MW_V2_T>(P, A, dense, cfg);
    case POT_GENERIC_T: return launch_co_kind<POT_GENERIC_T>(P, A, dense, cfg);
    default: return launch_co_kind<POT_GENERIC>(P, A, dense, cfg);
    }
}

} // namespace cw
