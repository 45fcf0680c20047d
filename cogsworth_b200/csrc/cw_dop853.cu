// cw_dop853.cu -- adaptive Dormand-Prince 8(5,3) kernel (sm_100a).
//
// Reference semantics (the test oracle restates the same algorithm on the CPU):
//   gala dop853.pyx dop853_integrate_hamiltonian: for every grid interval [t[j-1], t[j]] a fresh
//   Hairer dop853() with h = dt0 = t[1]-t[0] of the segment, itoler = 0, atol = rtol = 1e-10 by
//   default, safe = 0.9, fac1 = 0.333, fac2 = 6, beta = 0, hmax = xend - x, nmax = 100000,
//   nstiff = 1, uround = 2.3e-16; the error norm runs over the 6 components of the one orbit.
//   cogsworth restarts the integrator at every kick node (kicks.py:132-183).
//
//   DENSE = true is the dense-output variant (newer gala releases evaluate the
//   requested times with Hairer's continuous extension contd8 instead of stopping on every node):
//   ONE dop853() call per segment [kick node .. next kick node / last node], h0 = t[1]-t[0],
//   hmax = the whole segment, nmax attempts per segment.  Final-state runs never need the
//   interpolant (segments end exactly on the nodes whose state matters); store_all runs evaluate
//   it, after the three extra stages c14..c16, only for steps that contain a grid node.
//
// Execution model: persistent lanes, as in cw_leapfrog.cu, but the unit of work inside the loop is
// ONE RK STEP ATTEMPT of the lane's own current interval.  Lanes of a warp are not held at
// interval boundaries: a lane that needs sub-steps near the nucleus simply falls behind in ITS
// orbit while its neighbours keep advancing theirs (warp-level early exit: the warp leaves the
// loop when a ballot shows no lane has work and the queue is empty).  Divergence therefore costs
// only the difference in total attempts per orbit, not per interval.
//
// The first stage of every interval, k1 = f(t_j, y_j), is bit-identical to the derivative Hairer
// evaluates at the end of the previous accepted step (same y), so it is carried over instead of
// being recomputed: 12 gradient evaluations per accepted step instead of gala's 13.
//
// Stage algebra is kept in Hairer's first-order form (10 x 6 stage doubles, ~220 registers, 8 warps
// per SM) on purpose.  A Nystrom (second-order) folding that stores only the 3-vector accelerations
// was built and measured in round 1: 168 registers and 12 warps per SM, parity green, but 10-25 %
// SLOWER on every DOPRI853 workload.  These runs are bound by the single most expensive orbit
// (sequential by nature: 10^5 step attempts for a star orbiting inside 50 pc), i.e. by the latency
// of ONE lane, and six independent component chains per stage give that lane twice the
// instruction-level parallelism of three.
#include <cooperative_groups.h>
#include <climits>
#include "cw_kernels.cuh"
#include "cw_traj.cuh"
#include "../../include/cw_dop853_coeffs.h"

namespace cg = cooperative_groups;

namespace cw {

// The tableau lives in constant bank 3: a 64-bit literal costs two UMOVs at every use inside the unrolled
// stage algebra (220 of them), a constant-bank operand one uniform load.
__constant__ double k_A0201 = CW_A0201;
__constant__ double k_A0301 = CW_A0301;
__constant__ double k_A0302 = CW_A0302;
__constant__ double k_A0401 = CW_A0401;
__constant__ double k_A0403 = CW_A0403;
__constant__ double k_A0501 = CW_A0501;
__constant__ double k_A0503 = CW_A0503;
__constant__ double k_A0504 = CW_A0504;
__constant__ double k_A0601 = CW_A0601;
__constant__ double k_A0604 = CW_A0604;
__constant__ double k_A0605 = CW_A0605;
__constant__ double k_A0701 = CW_A0701;
__constant__ double k_A0704 = CW_A0704;
__constant__ double k_A0705 = CW_A0705;
__constant__ double k_A0706 = CW_A0706;
__constant__ double k_A0801 = CW_A0801;
__constant__ double k_A0804 = CW_A0804;
__constant__ double k_A0805 = CW_A0805;
__constant__ double k_A0806 = CW_A0806;
__constant__ double k_A0807 = CW_A0807;
__constant__ double k_A0901 = CW_A0901;
__constant__ double k_A0904 = CW_A0904;
__constant__ double k_A0905 = CW_A0905;
__constant__ double k_A0906 = CW_A0906;
__constant__ double k_A0907 = CW_A0907;
__constant__ double k_A0908 = CW_A0908;
__constant__ double k_A1001 = CW_A1001;
__constant__ double k_A1004 = CW_A1004;
__constant__ double k_A1005 = CW_A1005;
__constant__ double k_A1006 = CW_A1006;
__constant__ double k_A1007 = CW_A1007;
__constant__ double k_A1008 = CW_A1008;
__constant__ double k_A1009 = CW_A1009;
__constant__ double k_A1101 = CW_A1101;
__constant__ double k_A1104 = CW_A1104;
__constant__ double k_A1105 = CW_A1105;
__constant__ double k_A1106 = CW_A1106;
__constant__ double k_A1107 = CW_A1107;
__constant__ double k_A1108 = CW_A1108;
__constant__ double k_A1109 = CW_A1109;
__constant__ double k_A1110 = CW_A1110;
__constant__ double k_A1201 = CW_A1201;
__constant__ double k_A1204 = CW_A1204;
__constant__ double k_A1205 = CW_A1205;
__constant__ double k_A1206 = CW_A1206;
__constant__ double k_A1207 = CW_A1207;
__constant__ double k_A1208 = CW_A1208;
__constant__ double k_A1209 = CW_A1209;
__constant__ double k_A1210 = CW_A1210;
__constant__ double k_A1211 = CW_A1211;
__constant__ double k_B01 = CW_B01;
__constant__ double k_B06 = CW_B06;
__constant__ double k_B07 = CW_B07;
__constant__ double k_B08 = CW_B08;
__constant__ double k_B09 = CW_B09;
__constant__ double k_B10 = CW_B10;
__constant__ double k_B11 = CW_B11;
__constant__ double k_B12 = CW_B12;
__constant__ double k_BHH1 = CW_BHH1;
__constant__ double k_BHH2 = CW_BHH2;
__constant__ double k_BHH3 = CW_BHH3;
__constant__ double k_ER01 = CW_ER01;
__constant__ double k_ER06 = CW_ER06;
__constant__ double k_ER07 = CW_ER07;
__constant__ double k_ER08 = CW_ER08;
__constant__ double k_ER09 = CW_ER09;
__constant__ double k_ER10 = CW_ER10;
__constant__ double k_ER11 = CW_ER11;
__constant__ double k_ER12 = CW_ER12;
__constant__ double k_A1401 = CW_A1401;
__constant__ double k_A1407 = CW_A1407;
__constant__ double k_A1408 = CW_A1408;
__constant__ double k_A1409 = CW_A1409;
__constant__ double k_A1410 = CW_A1410;
__constant__ double k_A1411 = CW_A1411;
__constant__ double k_A1412 = CW_A1412;
__constant__ double k_A1413 = CW_A1413;
__constant__ double k_A1501 = CW_A1501;
__constant__ double k_A1506 = CW_A1506;
__constant__ double k_A1507 = CW_A1507;
__constant__ double k_A1508 = CW_A1508;
__constant__ double k_A1511 = CW_A1511;
__constant__ double k_A1512 = CW_A1512;
__constant__ double k_A1513 = CW_A1513;
__constant__ double k_A1514 = CW_A1514;
__constant__ double k_A1601 = CW_A1601;
__constant__ double k_A1606 = CW_A1606;
__constant__ double k_A1607 = CW_A1607;
__constant__ double k_A1608 = CW_A1608;
__constant__ double k_A1609 = CW_A1609;
__constant__ double k_A1613 = CW_A1613;
__constant__ double k_A1614 = CW_A1614;
__constant__ double k_A1615 = CW_A1615;
__constant__ double k_D401 = CW_D401;
__constant__ double k_D406 = CW_D406;
__constant__ double k_D407 = CW_D407;
__constant__ double k_D408 = CW_D408;
__constant__ double k_D409 = CW_D409;
__constant__ double k_D410 = CW_D410;
__constant__ double k_D411 = CW_D411;
__constant__ double k_D412 = CW_D412;
__constant__ double k_D413 = CW_D413;
__constant__ double k_D414 = CW_D414;
__constant__ double k_D415 = CW_D415;
__constant__ double k_D416 = CW_D416;
__constant__ double k_D501 = CW_D501;
__constant__ double k_D506 = CW_D506;
__constant__ double k_D507 = CW_D507;
__constant__ double k_D508 = CW_D508;
__constant__ double k_D509 = CW_D509;
__constant__ double k_D510 = CW_D510;
__constant__ double k_D511 = CW_D511;
__constant__ double k_D512 = CW_D512;
__constant__ double k_D513 = CW_D513;
__constant__ double k_D514 = CW_D514;
__constant__ double k_D515 = CW_D515;
__constant__ double k_D516 = CW_D516;
__constant__ double k_D601 = CW_D601;
__constant__ double k_D606 = CW_D606;
__constant__ double k_D607 = CW_D607;
__constant__ double k_D608 = CW_D608;
__constant__ double k_D609 = CW_D609;
__constant__ double k_D610 = CW_D610;
__constant__ double k_D611 = CW_D611;
__constant__ double k_D612 = CW_D612;
__constant__ double k_D613 = CW_D613;
__constant__ double k_D614 = CW_D614;
__constant__ double k_D615 = CW_D615;
__constant__ double k_D616 = CW_D616;
__constant__ double k_D701 = CW_D701;
__constant__ double k_D706 = CW_D706;
__constant__ double k_D707 = CW_D707;
__constant__ double k_D708 = CW_D708;
__constant__ double k_D709 = CW_D709;
__constant__ double k_D710 = CW_D710;
__constant__ double k_D711 = CW_D711;
__constant__ double k_D712 = CW_D712;
__constant__ double k_D713 = CW_D713;
__constant__ double k_D714 = CW_D714;
__constant__ double k_D715 = CW_D715;
__constant__ double k_D716 = CW_D716;

#ifndef CW_DP_BLOCK
#define CW_DP_BLOCK 128     // build knob (with CW_DP_MINB): threads per CTA
#endif
constexpr int DP_BLOCK = CW_DP_BLOCK;
#ifndef CW_DP_SPEC
#define CW_DP_SPEC 1       // build knob: evaluate the next step's first stage beside the step-size controller (see phase C)
#endif
#ifndef CW_DP_MINB
#define CW_DP_MINB 1       // resident CTAs per SM the kernel is compiled for (experiment knob; 1 = no register cap)
#endif

struct DpLane {
    double y[6];      // state at time x
    double k1[6];     // f(x, y) = (v, -grad)
    double x, xend;   // current time / end of the current interval (Myr)
    double h;         // trial step
    double hmax, facold, hlamb;
    double dt0;       // h0 of every interval of the current segment
    double tgrid;     // grid time of node j (grid unit, FP64-accumulated)
    double hgrid, t2, smyr;
    int64_t e, e1, oid;
    int j, L;
    int nstep, naccpt, iasti, nonsti;
    bool last, reject, append;
    // DENSE: the segment [j, next] is one dop853() call
    int next;         // node that ends the current segment (next kick node, or L)
    double tend;      // its grid time (grid unit)
    double tlast;     // grid time of the last accumulated node of the orbit (pre-pass)
};

__device__ __forceinline__ int64_t dp_queue_pop(unsigned long long *queue)
{
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(queue, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    return (int64_t)(base + g.thread_rank());
}

// f(x, y) = (v, -grad Phi(x, q)).  The separable kinds (Phi = f(t) Phi_0) keep the knot interval the orbit's clock is
// in as a cubic in registers (AmpSeg, as the leapfrog kernels do): the thirteen gradient evaluations of a step attempt
// then cost two comparisons and three FMAs for f(t) instead of a table look-up each.
template <int KIND>
__device__ __forceinline__ void fcn(const PotParams &P, LogTab ltab, AmpSeg &seg, double t, const double *y, double *f)
{
    double gx, gy, gz;
    if (KindTraits<KIND>::TS) {
        double Fx, Fy, Fz;
        force3<KindTraits<KIND>::BASE>(P, ltab, y[0], y[1], y[2], Fx, Fy, Fz);
        const double a = amp_seg_eval(P, seg, t);
        gx = (Fx * a) * y[0];
        gy = (Fy * a) * y[1];
        gz = (Fz * a) * y[2];
    } else {
        gradient<KIND>(P, ltab, t, y[0], y[1], y[2], gx, gy, gz);
    }
    f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
    f[3] = -gx; f[4] = -gy; f[5] = -gz;
}

__device__ __forceinline__ void dp_write_final(const IntegArgs &A, const DpLane &s, int status, int attempts = 0)
{
    const int64_t n = A.n;
    // diagnostics: the most step attempts any one orbit needed (its lane's sequential latency bounds the launch)
    if (A.stats && attempts > 0) atomicMax(A.stats + 5, (unsigned long long)attempts);
#pragma unroll
    for (int c = 0; c < 6; c++) A.w_final[c * n + s.oid] = s.y[c];
    if (status != CW_ORBIT_OK) {
        A.status[s.oid] = status;
        if (A.stats) atomicAdd(A.stats + 4, 1ull);      // cw_result.n_failed
    }
}

// kicks on node s.j: v += dv, the velocity half of k1 follows, the segment restarts with a new dt0
__device__ __forceinline__ void dp_apply_kicks(const IntegArgs &A, DpLane &s)
{
    while (s.e < s.e1 && A.ev_step[s.e] == s.j) {
        s.y[3] += A.ev_dv[3 * s.e + 0];
        s.y[4] += A.ev_dv[3 * s.e + 1];
        s.y[5] += A.ev_dv[3 * s.e + 2];
        const double tk = A.ev_tk[s.e];
        const double tn = (s.append && s.j + 1 == s.L) ? s.t2 : tk + s.hgrid;
        s.dt0 = seg_dt_myr(tk, tn, s.smyr);
        s.e++;
    }
    s.k1[0] = s.y[3]; s.k1[1] = s.y[4]; s.k1[2] = s.y[5];
}

// open the interval [node j, node j+1] (DENSE: the segment [node j, next sync node]) = the argument
// set of one dop853() call
template <bool DENSE>
__device__ __forceinline__ void dp_begin_interval(const IntegArgs &A, DpLane &s)
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

// PILOT = true: integrate only the first A.pilot_intervals intervals of every orbit and write a cost
// estimate (attempts per interval x number of intervals) to A.cost_key, nothing else.  The host
// sorts the queue by that key so the expensive (short-period, sub-stepping) orbits start first:
// with one sequential orbit per lane the makespan is max(longest orbit, total work / lanes), and a
// long orbit started late would leave the whole GPU waiting for a single lane.
template <int KIND, bool STORE, bool PILOT, bool DENSE>
__global__ void __launch_bounds__(DP_BLOCK, CW_DP_MINB) dop853_kernel(const __grid_constant__ PotParams P,
                                                          const __grid_constant__ IntegArgs A)
{
    const unsigned FULL = 0xffffffffu;
    __shared__ double2 ltab_mem[LOG_TABLE_N];
    extern __shared__ __align__(16) unsigned char traj_smem[];
    const LogTab ltab = log_table_init(ltab_mem);
    TrajWriterDop tw;
    if (STORE && !PILOT) tw.init(traj_smem, A);
    int64_t col_base = 0;
    DpLane s;
    AmpSeg seg;                        // separable kinds only (dead code otherwise)
    amp_seg_reset(seg);
#pragma unroll
    for (int c = 0; c < 6; c++) { s.y[c] = (c == 0) ? 8.0 : 0.0; s.k1[c] = 0.0; }
    s.x = 0.0; s.xend = 1.0; s.h = 1.0; s.hmax = 1.0; s.facold = 1e-4; s.hlamb = 0.0; s.dt0 = 1.0;
    s.tgrid = 0.0; s.hgrid = 1.0; s.t2 = 0.0; s.smyr = 1.0;
    s.e = s.e1 = 0; s.oid = -1; s.j = 0; s.L = 0;
    s.nstep = s.naccpt = s.iasti = s.nonsti = 0;
    s.last = s.reject = s.append = false;
    s.next = 0; s.tend = 0.0; s.tlast = 0.0;
    // DENSE && STORE: continuous extension of the last accepted step (Hairer rcont1..8), valid while `pend`
    double rc[8][6];
    double dn_xold = 0.0, dn_h = 1.0;
    bool pend = false;
    bool active = false, exhausted = false, interval_done = false;
    int attempts = 0, Lstop = 0;       // PILOT bookkeeping
    unsigned long long n_acc = 0, n_rej = 0, n_fcn = 0, n_int = 0;
    const int64_t nmax = (A.nmax > 0) ? A.nmax : 100000;
    const double atol = A.atol, rtol = A.rtol;

    for (;;) {
        // ---- phase A0 (DENSE store_all): grid nodes strictly inside the segment that the last accepted
        // step passed over are evaluated with contd8, one node per lane per warp-uniform trip
        if (DENSE && STORE && !PILOT) {
            for (;;) {
                const double tnx = (s.tgrid + s.hgrid) * s.smyr;   // next grid node; the grid may run backward in time
                const bool emit = active && pend && (s.j + 1 < s.next) && ((s.hgrid > 0.0) ? (tnx <= s.x) : (tnx >= s.x));
                if (!__any_sync(FULL, emit)) break;
                if (emit) {
                    s.tgrid += s.hgrid;
                    s.j += 1;
                    n_int++;
                    const double th = (s.tgrid * s.smyr - dn_xold) / dn_h, th1 = 1.0 - th;
                    double v[6];
#pragma unroll
                    for (int i = 0; i < 6; i++)
                        v[i] = fma(th, fma(th1, fma(th, fma(th1, fma(th, fma(th1, fma(th, rc[7][i], rc[6][i]), rc[5][i]),
                                   rc[4][i]), rc[3][i]), rc[2][i]), rc[1][i]), rc[0][i]);
                    tw.put(col_base + s.j, v[0], v[1], v[2], v[3], v[4], v[5], false);
                }
                tw.flush_warp(A);
            }
            pend = false;
        }
        // ---- phase A: lanes that completed an interval stand on node j+1 (DENSE: on the segment's end node)
        if (active && interval_done) {
            interval_done = false;
            if (DENSE) {
                n_int += (unsigned long long)(s.next - s.j);
                s.j = s.next;
                s.tgrid = s.tend;
            } else {
                s.j += 1;
                s.tgrid = (s.append && s.j == s.L) ? s.t2 : s.tgrid + s.hgrid;
                n_int++;
            }
            dp_apply_kicks(A, s);
            if (STORE && !PILOT)
                tw.put(col_base + s.j, s.y[0], s.y[1], s.y[2], s.y[3], s.y[4], s.y[5], s.j == s.L);
            if (PILOT && s.j == Lstop) {
                A.cost_key[s.oid] = __float_as_int((float)attempts * ((float)s.L / (float)Lstop));
                active = false;
            } else if (s.j == s.L) {
                dp_write_final(A, s, CW_ORBIT_OK, attempts);
                active = false;
            } else {
                dp_begin_interval<DENSE>(A, s);
            }
        }
        if (STORE && !PILOT) tw.flush_warp(A);
        // ---- phase B: idle lanes pull the next orbit
        while (!active && !exhausted) {
            const int64_t p = dp_queue_pop(A.queue);
            if (p >= A.n) { exhausted = true; break; }
            const int64_t o = A.order ? (int64_t)A.order[p] : p;
            const int64_t n = A.n;
            s.oid = o;
#pragma unroll
            for (int c = 0; c < 6; c++) s.y[c] = A.w0[c * n + o];
            s.j = 0;
            s.L = A.n_nodes[o] - 1;
            if (A.status[o] != CW_ORBIT_OK || s.L < 0) {
                if (PILOT) A.cost_key[o] = 0;
                else dp_write_final(A, s, CW_ORBIT_OK); // status already set by the pre-pass
                continue;
            }
            attempts = 0;
            Lstop = min(s.L, A.pilot_intervals);
            s.hgrid = A.dt[o]; s.t2 = A.t2[o]; s.smyr = A.to_myr[o];
            s.append = (A.flags[o] & CW_GRID_APPEND_T2) != 0;
            s.tgrid = A.t1[o];
            if (DENSE) s.tlast = A.t_last[o];
            s.e = s.e1 = 0;
            if (A.ev_off) { s.e = A.ev_off[o] - A.ev_base; s.e1 = A.ev_off[o + 1] - A.ev_base; }
            {
                const double tn = (s.append && s.L == 1) ? s.t2 : s.tgrid + s.hgrid;
                s.dt0 = seg_dt_myr(s.tgrid, tn, s.smyr);
            }
            dp_apply_kicks(A, s);
            if (STORE && !PILOT) {
                col_base = A.traj_off[o] - A.traj_base;
                if (s.L == 0) traj_store_direct(A, col_base, s.y[0], s.y[1], s.y[2], s.y[3], s.y[4], s.y[5]);
                else tw.put(col_base, s.y[0], s.y[1], s.y[2], s.y[3], s.y[4], s.y[5], false);
            }
            if (s.L == 0) {
                if (PILOT) A.cost_key[o] = 0;
                else dp_write_final(A, s, CW_ORBIT_OK);
                continue;
            }
            fcn<KIND>(P, ltab, seg, s.tgrid * s.smyr, s.y, s.k1);
            n_fcn++;
            dp_begin_interval<DENSE>(A, s);
            active = true;
        }
        if (STORE && !PILOT) tw.flush_warp(A);
        if (!__any_sync(FULL, active)) break;   // warp-level early exit

        // ---- phase C: one step attempt of dopcor() for every lane that has an open interval
        if (active) {
            int fail = 0;
            if (s.nstep > nmax) fail = CW_ORBIT_NMAX;
            else if (0.1 * fabs(s.h) <= fabs(s.x) * 2.3e-16) fail = CW_ORBIT_STEP_TOO_SMALL;
            if (fail) {
                if (PILOT) A.cost_key[s.oid] = 0x7f000000;   // failing orbits are the most expensive: first
                else dp_write_final(A, s, fail);
                active = false;
                continue;
            }
            attempts++;
            const double posneg = (s.xend - s.x > 0.0) ? 1.0 : -1.0;
            if (__dmul_rn(__dsub_rn(__dadd_rn(s.x, __dmul_rn(1.01, s.h)), s.xend), posneg) > 0.0) { s.h = __dsub_rn(s.xend, s.x); s.last = true; }
            s.nstep++;
            const double h = s.h;
            const double *y = s.y, *k1 = s.k1;
            double k2[6], k3[6], k4[6], k5[6], k6[6], k7[6], k8[6], k9[6], k10[6], yy1[6];

#define CW_STAGE(EXPR) _Pragma("unroll") for (int i = 0; i < 6; i++) yy1[i] = fma(h, (EXPR), y[i]);
            CW_STAGE(k_A0201 * k1[i])
            fcn<KIND>(P, ltab, seg, fma(CW_C02, h, s.x), yy1, k2);
            CW_STAGE(fma(k_A0302, k2[i], k_A0301 * k1[i]))
            fcn<KIND>(P, ltab, seg, fma(CW_C03, h, s.x), yy1, k3);
            CW_STAGE(fma(k_A0403, k3[i], k_A0401 * k1[i]))
            fcn<KIND>(P, ltab, seg, fma(CW_C04, h, s.x), yy1, k4);
            CW_STAGE(fma(k_A0504, k4[i], fma(k_A0503, k3[i], k_A0501 * k1[i])))
            fcn<KIND>(P, ltab, seg, fma(CW_C05, h, s.x), yy1, k5);
            CW_STAGE(fma(k_A0605, k5[i], fma(k_A0604, k4[i], k_A0601 * k1[i])))
            fcn<KIND>(P, ltab, seg, fma(CW_C06, h, s.x), yy1, k6);
            CW_STAGE(fma(k_A0706, k6[i], fma(k_A0705, k5[i], fma(k_A0704, k4[i], k_A0701 * k1[i]))))
            fcn<KIND>(P, ltab, seg, fma(CW_C07, h, s.x), yy1, k7);
            CW_STAGE(fma(k_A0807, k7[i], fma(k_A0806, k6[i], fma(k_A0805, k5[i], fma(k_A0804, k4[i], k_A0801 * k1[i])))))
            fcn<KIND>(P, ltab, seg, fma(CW_C08, h, s.x), yy1, k8);
            CW_STAGE(fma(k_A0908, k8[i], fma(k_A0907, k7[i], fma(k_A0906, k6[i], fma(k_A0905, k5[i], fma(k_A0904, k4[i], k_A0901 * k1[i]))))))
            fcn<KIND>(P, ltab, seg, fma(CW_C09, h, s.x), yy1, k9);
            CW_STAGE(fma(k_A1009, k9[i], fma(k_A1008, k8[i], fma(k_A1007, k7[i], fma(k_A1006, k6[i], fma(k_A1005, k5[i], fma(k_A1004, k4[i], k_A1001 * k1[i])))))))
            fcn<KIND>(P, ltab, seg, fma(CW_C10, h, s.x), yy1, k10);
            CW_STAGE(fma(k_A1110, k10[i], fma(k_A1109, k9[i], fma(k_A1108, k8[i], fma(k_A1107, k7[i], fma(k_A1106, k6[i], fma(k_A1105, k5[i], fma(k_A1104, k4[i], k_A1101 * k1[i]))))))))
            fcn<KIND>(P, ltab, seg, fma(CW_C11, h, s.x), yy1, k2);
            const double xph = __dadd_rn(s.x, h);
            CW_STAGE(fma(k_A1211, k2[i], fma(k_A1210, k10[i], fma(k_A1209, k9[i], fma(k_A1208, k8[i], fma(k_A1207, k7[i], fma(k_A1206, k6[i], fma(k_A1205, k5[i], fma(k_A1204, k4[i], k_A1201 * k1[i])))))))))
            fcn<KIND>(P, ltab, seg, xph, yy1, k3);
#undef CW_STAGE
            n_fcn += 11;

            double err = 0.0, err2 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                k4[i] = fma(k_B12, k3[i], fma(k_B11, k2[i], fma(k_B10, k10[i], fma(k_B09, k9[i],
                        fma(k_B08, k8[i], fma(k_B07, k7[i], fma(k_B06, k6[i], k_B01 * k1[i])))))));
                k5[i] = fma(h, k4[i], y[i]);
                const double sk = fma(rtol, fmax(fabs(y[i]), fabs(k5[i])), atol);
                const double isk = rcp_fast(sk);
                double erri = k4[i] - k_BHH1 * k1[i] - k_BHH2 * k9[i] - k_BHH3 * k3[i];
                double sq = erri * isk;
                err2 = fma(sq, sq, err2);
                erri = fma(k_ER12, k3[i], fma(k_ER11, k2[i], fma(k_ER10, k10[i], fma(k_ER09, k9[i],
                       fma(k_ER08, k8[i], fma(k_ER07, k7[i], fma(k_ER06, k6[i], k_ER01 * k1[i])))))));
                sq = erri * isk;
                err = fma(sq, sq, err);
            }
#if CW_DP_SPEC
            // f(x + h, y_new), the first stage of the next step, is needed only if this attempt is accepted -- 99.6 % are --
            // and depends on nothing the controller computes: evaluated HERE, its long dependent chain runs beside the
            // controller's (four reciprocal square roots in a row), instead of after it.  A rejected attempt discards it.
            fcn<KIND>(P, ltab, seg, xph, k5, k4);
#endif
            // Controller arithmetic on the DFMA pipe with the kernels' own branch-free rsqrt / reciprocal (the
            // library sqrt / divide carry slow-path branches and ~4x the instructions; a lane's step attempts
            // are strictly sequential, so every instruction here is on the critical path of the slowest orbit).
            double deno = fma(0.01, err2, err);
            if (!(deno > 1.0e-290)) deno = 1.0;            // Hairer: deno <= 0 -> 1 (and keeps the rsqrt seed normal)
            err = fabs(h) * err * rsqrt_fast(deno * 6.0);
            if (!(err <= 1.79769313486231571e+308)) {   // NaN or inf: Hairer would spin to nmax
                if (PILOT) A.cost_key[s.oid] = 0;
                else dp_write_final(A, s, CW_ORBIT_NONFINITE);
                active = false;
                continue;
            }
            // err^(1/8) (beta = 0) by three square roots; below 1e-280 the factor is clamped to facc2 anyway
            const double e0 = fmax(err, 1.0e-280);
            const double e1 = e0 * rsqrt_fast(e0);
            const double e2 = e1 * rsqrt_fast(e1);
            const double fac11 = e2 * rsqrt_fast(e2);
            const double facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0, inv_safe = 1.0 / 0.9;
            const double fac = fmax(facc2, fmin(facc1, fac11 * inv_safe));
            double hnew = h * rcp_fast(fac);
            if (err <= 1.0) {
                // ---- step accepted
                s.facold = fmax(err, 1.0e-4);
                s.naccpt++;
                n_acc++;
#if !CW_DP_SPEC
                fcn<KIND>(P, ltab, seg, xph, k5, k4);  // first stage of the next step
#endif
                n_fcn++;
                // stiffness detection (nstiff = 1: every accepted step)
                double stnum = 0.0, stden = 0.0;
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double a = k4[i] - k3[i];
                    stnum = fma(a, a, stnum);
                    const double b = k5[i] - yy1[i];
                    stden = fma(b, b, stden);
                }
                // hlamb = h sqrt(stnum / stden) > 6.1  <=>  h > 0 and h^2 stnum > 37.21 stden  (no root, no divide):
                // Hairer's hlamb carries the sign of h, so a backward integration (h < 0, hydro/rewind.py) never
                // trips the test; hlamb keeps its previous verdict when stden == 0, as in dopcor
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
                if (DENSE && STORE && !PILOT) {
                    // a grid node of this segment inside (x, x+h]: dop853.c "final preparation for dense output"
                    const double tnx = (s.tgrid + s.hgrid) * s.smyr;
                    if ((s.j + 1 < s.next) && ((s.hgrid > 0.0) ? (tnx <= xph) : (tnx >= xph))) {
                        double ka[6], kb[6], kc[6];
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            const double ydiff = k5[i] - y[i];
                            const double bspl = fma(h, k1[i], -ydiff);
                            rc[0][i] = y[i];
                            rc[1][i] = ydiff;
                            rc[2][i] = bspl;
                            rc[3][i] = ydiff - h * k4[i] - bspl;
                            rc[4][i] = fma(k_D412, k3[i], fma(k_D411, k2[i], fma(k_D410, k10[i], fma(k_D409, k9[i], fma(k_D408, k8[i], fma(k_D407, k7[i], fma(k_D406, k6[i], k_D401 * k1[i])))))));
                            rc[5][i] = fma(k_D512, k3[i], fma(k_D511, k2[i], fma(k_D510, k10[i], fma(k_D509, k9[i], fma(k_D508, k8[i], fma(k_D507, k7[i], fma(k_D506, k6[i], k_D501 * k1[i])))))));
                            rc[6][i] = fma(k_D612, k3[i], fma(k_D611, k2[i], fma(k_D610, k10[i], fma(k_D609, k9[i], fma(k_D608, k8[i], fma(k_D607, k7[i], fma(k_D606, k6[i], k_D601 * k1[i])))))));
                            rc[7][i] = fma(k_D712, k3[i], fma(k_D711, k2[i], fma(k_D710, k10[i], fma(k_D709, k9[i], fma(k_D708, k8[i], fma(k_D707, k7[i], fma(k_D706, k6[i], k_D701 * k1[i])))))));
                        }
                        double yd[6];
#pragma unroll
                        for (int i = 0; i < 6; i++)
                            yd[i] = fma(h, fma(k_A1413, k4[i], fma(k_A1412, k3[i], fma(k_A1411, k2[i], fma(k_A1410, k10[i], fma(k_A1409, k9[i], fma(k_A1408, k8[i], fma(k_A1407, k7[i], k_A1401 * k1[i]))))))), y[i]);
                        fcn<KIND>(P, ltab, seg, fma(0.1, h, s.x), yd, ka);
#pragma unroll
                        for (int i = 0; i < 6; i++)
                            yd[i] = fma(h, fma(k_A1514, ka[i], fma(k_A1513, k4[i], fma(k_A1512, k3[i], fma(k_A1511, k2[i], fma(k_A1508, k8[i], fma(k_A1507, k7[i], fma(k_A1506, k6[i], k_A1501 * k1[i]))))))), y[i]);
                        fcn<KIND>(P, ltab, seg, fma(0.2, h, s.x), yd, kb);
#pragma unroll
                        for (int i = 0; i < 6; i++)
                            yd[i] = fma(h, fma(k_A1615, kb[i], fma(k_A1614, ka[i], fma(k_A1613, k4[i], fma(k_A1609, k9[i], fma(k_A1608, k8[i], fma(k_A1607, k7[i], fma(k_A1606, k6[i], k_A1601 * k1[i]))))))), y[i]);
                        fcn<KIND>(P, ltab, seg, fma(7.0 / 9.0, h, s.x), yd, kc);
                        n_fcn += 3;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            rc[4][i] = h * fma(k_D416, kc[i], fma(k_D415, kb[i], fma(k_D414, ka[i], fma(k_D413, k4[i], rc[4][i]))));
                            rc[5][i] = h * fma(k_D516, kc[i], fma(k_D515, kb[i], fma(k_D514, ka[i], fma(k_D513, k4[i], rc[5][i]))));
                            rc[6][i] = h * fma(k_D616, kc[i], fma(k_D615, kb[i], fma(k_D614, ka[i], fma(k_D613, k4[i], rc[6][i]))));
                            rc[7][i] = h * fma(k_D716, kc[i], fma(k_D715, kb[i], fma(k_D714, ka[i], fma(k_D713, k4[i], rc[7][i]))));
                        }
                        dn_xold = s.x;
                        dn_h = h;
                        pend = true;
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; i++) { s.k1[i] = k4[i]; s.y[i] = k5[i]; }
                s.x = xph;
                if (stiff) {
                    if (PILOT) A.cost_key[s.oid] = 0;
                    else dp_write_final(A, s, CW_ORBIT_STIFF);
                    active = false;
                    continue;
                }
                if (s.last) {
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
                if (s.naccpt >= 1) n_rej++;
                s.last = false;
                s.h = hnew;
            }
        }
    }
    if (A.stats && !PILOT) {
        for (int off = 16; off > 0; off >>= 1) {
            n_acc += __shfl_down_sync(FULL, n_acc, off);
            n_rej += __shfl_down_sync(FULL, n_rej, off);
            n_fcn += __shfl_down_sync(FULL, n_fcn, off);
            n_int += __shfl_down_sync(FULL, n_int, off);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_acc) atomicAdd(A.stats + 0, n_acc);
            if (n_rej) atomicAdd(A.stats + 1, n_rej);
            if (n_fcn) atomicAdd(A.stats + 2, n_fcn);
            if (n_int) atomicAdd(A.stats + 3, n_int);
        }
    }
}

template <int KIND, bool STORE, bool PILOT, bool DENSE>
static cudaError_t launch_dp(const PotParams &P, const IntegArgs &A, const LaunchCfg &cfg)
{
    int per_sm = 0;
    const size_t dyn = (STORE && !PILOT) ? (size_t)DP_BLOCK * TrajWriterDop::LANE_BYTES : 0;
    cudaError_t err = cudaSuccess;
    if (dyn) {
        err = cudaFuncSetAttribute(dop853_kernel<KIND, STORE, PILOT, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (err != cudaSuccess) return err;
    }
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dop853_kernel<KIND, STORE, PILOT, DENSE>, DP_BLOCK, dyn);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)cfg.sm_count * per_sm;
    const int64_t need = (A.n + DP_BLOCK - 1) / DP_BLOCK;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    dop853_kernel<KIND, STORE, PILOT, DENSE><<<(unsigned)grid, DP_BLOCK, dyn, cfg.stream>>>(P, A);
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_dp_kind(const PotParams &P, const IntegArgs &A, bool store_all, bool dense, const LaunchCfg &cfg)
{
    if (dense) return store_all ? launch_dp<KIND, true, false, true>(P, A, cfg) : launch_dp<KIND, false, false, true>(P, A, cfg);
    return store_all ? launch_dp<KIND, true, false, false>(P, A, cfg) : launch_dp<KIND, false, false, false>(P, A, cfg);
}

cudaError_t launch_dop853(const PotParams &P, const IntegArgs &A, bool store_all, bool dense, const LaunchCfg &cfg)
{
    if (A.n <= 0) return cudaSuccess;
    if (dense && !A.t_last) return cudaErrorInvalidValue;
    switch (P.kind) {
    case POT_MW_V1: return launch_dp_kind<POT_MW_V1>(P, A, store_all, dense, cfg);
    case POT_MW_V2: return launch_dp_kind<POT_MW_V2>(P, A, store_all, dense, cfg);
    case POT_GENERAL: return launch_dp_kind<POT_GENERAL>(P, A, store_all, dense, cfg);
    case POT_ROTATED: return launch_dp_kind<POT_ROTATED>(P, A, store_all, dense, cfg);
    case POT_LM10: return launch_dp_kind<POT_LM10>(P, A, store_all, dense, cfg);
    case POT_MW_V2_T: return launch_dp_kind<POT_MW_V2_T>(P, A, store_all, dense, cfg);
    case POT_GENERIC_T: return launch_dp_kind<POT_GENERIC_T>(P, A, store_all, dense, cfg);
    default: return launch_dp_kind<POT_GENERIC>(P, A, store_all, dense, cfg);
    }
}

// The pilot always runs the restart-per-interval form: attempts per grid interval is the cost proxy of
// both modes (it resolves exactly the expensive, sub-stepping orbits that decide the makespan).
cudaError_t launch_dop853_pilot(const PotParams &P, const IntegArgs &A, const LaunchCfg &cfg)
{
    if (A.n <= 0) return cudaSuccess;
    switch (P.kind) {
    case POT_MW_V1: return launch_dp<POT_MW_V1, false, true, false>(P, A, cfg);
    case POT_MW_V2: return launch_dp<POT_MW_V2, false, true, false>(P, A, cfg);
    case POT_GENERAL: return launch_dp<POT_GENERAL, false, true, false>(P, A, cfg);
    case POT_ROTATED: return launch_dp<POT_ROTATED, false, true, false>(P, A, cfg);
    case POT_LM10: return launch_dp<POT_LM10, false, true, false>(P, A, cfg);
    case POT_MW_V2_T: return launch_dp<POT_MW_V2_T, false, true, false>(P, A, cfg);
    case POT_GENERIC_T: return launch_dp<POT_GENERIC_T, false, true, false>(P, A, cfg);
    default: return launch_dp<POT_GENERIC, false, true, false>(P, A, cfg);
    }
}

} // namespace cw
