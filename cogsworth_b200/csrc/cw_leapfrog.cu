// cw_leapfrog.cu -- grid pre-pass and the fixed-step leapfrog kernel (sm_100a).
//
// Reference semantics (restated in oracle/cw_oracle.c):
//   time grid + kick node      cogsworth kicks.py:97,118-122,134; gala timespec.py
//   leapfrog                   gala leapfrog.pyx c_init_velocity / c_leapfrog_step
//   segment restart at kicks   cogsworth kicks.py:132-183
//
// Execution model: PERSISTENT LANES.  The grid is sized to fill the 148 SMs exactly once; every
// lane (thread) owns one orbit at a time and pulls the next one from a global queue (longest
// orbits first) the moment its own finishes, so divergent step counts never leave lanes idle
// behind a warp-wide barrier and there is no wave-quantisation tail.  Inside a warp the hot loop
// runs a warp-uniform number of branch-free steps (the minimum over lanes of the distance to the
// lane's next kick / last node), then the lanes that reached such a node handle it.
#include <cooperative_groups.h>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include "cw_kernels.cuh"
#include "cw_traj.cuh"

namespace cg = cooperative_groups;

namespace cw {

// ------------------------------------------------------------------------------------------
// Grid pre-pass: one thread per orbit walks T_{j+1} = T_j + dt in FP64 exactly as
// parse_time_specification does and resolves, for every event, k = max(cursor, last j with
// T_j < t_event) -- the bit-exact kick timing the reference defines at kicks.py:134.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_prepass_kernel(const __grid_constant__ IntegArgs A, int32_t *iota,
                                                           int32_t *sort_key)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const double a = A.t1[i], b = A.t2[i], h = A.dt[i];
    const bool append = (A.flags[i] & CW_GRID_APPEND_T2) != 0;
    int64_t e = 0, e1 = 0;
    if (A.ev_off) {
        e = A.ev_off[i] - A.ev_base;
        e1 = A.ev_off[i + 1] - A.ev_base;
    }
    int st = CW_ORBIT_OK;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double te = (e < e1) ? A.ev_time[e] : INF;
    double te_prev = -INF;
    int j = 0, cursor = 0;
    double T = a, Tprev = a;
    if (te != te) st = CW_ORBIT_INCONSISTENT;      // NaN event time
    // One loop, one FP64 compare + one DADD per node on the fast path; the rare branch (an event
    // lands, or the grid ends) sits INSIDE the loop body so the warp reconverges every iteration.
    double lim = fmin(b, te);
    const bool backward = (h < 0.0) && (b < a);    // rewind (cogsworth hydro/rewind.py:16-19): t2 < t1, dt < 0
    if (!(h > 0.0) && !backward) st = CW_ORBIT_INCONSISTENT;    // gala raises for dt <= 0 with t2 > t1
    bool done = false;
    if (backward) {
        // gala's backward grid: nodes while T > t2 (same sequential FP64 accumulation); events never occur on
        // this branch (kicks.py only integrates forward)
        while (T > b && j < 1000000) {
            Tprev = T;
            j++;
            T += h;
        }
        if (e < e1) st = CW_ORBIT_INCONSISTENT;
        done = true;
    }
    while (!done) {
        // 8 nodes per trip while none of them reaches `lim`: the same sequential FP64 additions, but
        // one compare and one branch for eight of them (the add chain is latency-, not pipe-bound).
        // Both arms sit inside one if/else so the warp reconverges at the bottom of every trip.
        const double a1 = T + h, a2 = a1 + h, a3 = a2 + h, a4 = a3 + h;
        const double a5 = a4 + h, a6 = a5 + h, a7 = a6 + h, a8 = a7 + h;
        if (a7 < lim && j + 8 <= 1000000 && h > 0.0) {
            Tprev = a7;
            T = a8;
            j += 8;
        } else {
            if (!(T < lim) || j >= 1000000) {
                if (!(T < b) || j >= 1000000) {
                    done = true;                   // the grid ends here
                } else {
                    // node j exists at time T and every pending event with t_event <= T lands before it
                    while (!(T < te)) {
                        const int k = max(cursor, j - 1);
                        A.ev_step[e] = k;
                        A.ev_tk[e] = Tprev;
                        cursor = k;
                        if (te < te_prev) st = CW_ORBIT_INCONSISTENT; // events must be chronological
                        te_prev = te;
                        e++;
                        te = (e < e1) ? A.ev_time[e] : INF;
                        if (te != te) st = CW_ORBIT_INCONSISTENT;
                    }
                    lim = fmin(b, te);
                }
            }
            if (!done) {
                Tprev = T;
                j++;
                T += h;
            }
        }
    }
    const int n_acc = j;
    // forward: t2 is appended after the last node < t2; backward: gala appends it unless the last node IS t2
    const int n_nodes = n_acc + ((append && n_acc > 0 && !(backward && Tprev == b)) ? 1 : 0);
    // events not before any accumulated node boundary: after the last accumulated node
    while (e < e1) {
        int k;
        if (append && te > b) k = n_nodes - 1;        // later than t2 itself: lands on the final node
        else k = max(cursor, n_acc - 1);
        A.ev_step[e] = k;
        A.ev_tk[e] = Tprev;
        cursor = k;
        if (te < te_prev) st = CW_ORBIT_INCONSISTENT;
        te_prev = te;
        e++;
        te = (e < e1) ? A.ev_time[e] : INF;
    }
    // a kick on the final node leaves no tail segment: the reference cannot stitch such an orbit
    if (n_nodes <= 0 || (cursor >= n_nodes - 1 && (e1 > (A.ev_off ? A.ev_off[i] - A.ev_base : 0))))
        st = CW_ORBIT_INCONSISTENT;
    // store_all: the caller's trajectory slot must hold every node (a too-small slot would overwrite the next orbit)
    if (A.traj_off && st == CW_ORBIT_OK && A.traj_off[i + 1] - A.traj_off[i] < (int64_t)n_nodes) st = CW_ORBIT_CAPACITY;
    A.n_nodes[i] = n_nodes;
    A.status[i] = st;
    if (st != CW_ORBIT_OK && A.stats) atomicAdd(A.stats + 4, 1ull);     // cw_result.n_failed
    if (A.t_last) A.t_last[i] = Tprev;     // time of node n_acc - 1 (DOP853 dense mode: end of the last segment)
    if (iota) iota[i] = (int32_t)i;
    if (sort_key) {
        // queue order: longest orbit first; among equal lengths the (stable) sort keeps index order, so the
        // 32 lanes of a warp write 32 NEIGHBOURING trajectories.  Measured on store_all: lanes whose streams
        // are >= 32 KB apart (orbits grouped by column phase instead) reach 2.3 TB/s, neighbours 3.4 TB/s;
        // phase alignment is the warp clock's job (cw_leapfrog.cu), not the queue's.
        int key = min(max(n_nodes, 0), (1 << 27) - 1) << 4;
        sort_key[i] = key;
    }
}

cudaError_t launch_grid_prepass(const IntegArgs &A, int32_t *iota_out, int32_t *sort_key_out, const LaunchCfg &cfg)
{
    if (A.n <= 0) return cudaSuccess;
    const int block = 256;
    const int64_t grid = (A.n + block - 1) / block;
    grid_prepass_kernel<<<(unsigned)grid, block, 0, cfg.stream>>>(A, iota_out, sort_key_out);
    return cudaGetLastError();
}

// Orbit.t of stored trajectories: the grid itself, converted to Myr (kicks.py:188).  t is a pure
// function of (t1, dt, t2, flags), so it is written by this side kernel (thread per orbit, same FP64
// accumulation as the pre-pass) instead of riding through the integrator kernels.
__global__ void __launch_bounds__(256) traj_times_kernel(const __grid_constant__ IntegArgs A)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n || A.status[i] != CW_ORBIT_OK) return;
    const double b = A.t2[i], h = A.dt[i], sm = A.to_myr[i];
    const int n = A.n_nodes[i];
    const bool append = (A.flags[i] & CW_GRID_APPEND_T2) != 0;
    double *out = A.traj_t + (A.traj_off[i] - A.traj_base);
    double T = A.t1[i];
    for (int j = 0; j < n; j++) {
        out[j] = ((append && j == n - 1) ? b : T) * sm;
        T += h;
    }
}

cudaError_t launch_traj_times(const IntegArgs &A, const LaunchCfg &cfg)
{
    if (A.n <= 0 || !A.traj_t || !A.traj_off) return cudaSuccess;
    const int block = 256;
    traj_times_kernel<<<(unsigned)((A.n + block - 1) / block), block, 0, cfg.stream>>>(A);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Leapfrog
// ------------------------------------------------------------------------------------------
#ifndef CW_LF_BLOCK
#define CW_LF_BLOCK 256    // experiment knob (with CW_LF_REGS): threads per CTA of the final-state kernel
#endif
constexpr int LF_BLOCK = CW_LF_BLOCK; // final-state kernel: 4 CTAs x 256 threads per SM (64 registers)
// store_all kernel: ONE CTA of 512 threads per SM (cw_traj.cuh: the windows of 256 lanes fill the shared memory, 776 B
// per lane = 196 KB; the other 256 lanes stage in tensor memory, and a kernel that allocates tensor memory is given one
// CTA per SM by the occupancy calculator whatever it asks for, so the 16 warps live in one CTA that owns all 512 TMEM
// columns); 128 registers per thread.
constexpr int LF_BLOCK_STORE = TRAJ_STORE_BLOCK;   // 512 (256 without TRAJ_TSTAGE; cw_traj.cuh)
#ifndef CW_LF_REGS
#define CW_LF_REGS 64      // experiment knob: 80 -> 3 resident CTAs (measured: see DESIGN.md)
#endif
// Tensor memory as a second staging level (a completed lower half-line window parked until its upper half arrives):
// what made 64-byte runs leave as full 128-byte lines with 8-sample windows (round 1).  With 16-sample windows every
// run IS a full line, and parking only adds shared-memory traffic (measured 3.84 vs 4.20 TB/s): on for TRAJ_S = 8 only.
#ifndef CW_TRAJ_TMEM
#define CW_TRAJ_TMEM (CW_TRAJ_S == 8)
#endif
#ifndef CW_LF_STORE_REGS
#define CW_LF_STORE_REGS 168   // trajectory kernel with 256 lanes per SM
#endif
#ifndef CW_LF_GEN_REGS
#define CW_LF_GEN_REGS 80  // final-state kernel of the general and separable kinds: 3 resident CTAs (measured +14 % over 128 on
                           // LM10 in the general kernel, +2 % over 64 in the evolving Milky Way); the static rotated kinds
                           // (POT_ROTATED, POT_LM10) take the 64 registers of the Milky Way kernels (+4.5 % over 80 on LM10)
#endif
// Register budgets (__maxnreg__ on the kernel): 64 for the final-state kernel = 4 resident CTAs of 256 threads,
// 128 for the store_all kernel = its single 512-thread CTA.

struct Lane {
    double x, y, z, vx, vy, vz;     // state at node j (v synchronised only at sync nodes)
    double vhx, vhy, vhz;           // half-step velocity
    double Fxy, Fy, Fz;             // force factors at node j: grad = (Fxy x, Fy y, Fz z); Fy = Fxy unless the generic
                                    // composite holds a component flattened along x or y
    double dt;                      // leapfrog dt of the current segment (Myr)
    double hgrid, t2, smyr;         // grid step, end time (grid unit), grid unit -> Myr
    double tcur;                    // STORE only: grid time of node j
    int64_t e, e1;                  // pending events
    int64_t oid;                    // orbit id within the shard
    int j, next, L;                 // current node, next sync node, last node
    bool append;
    AmpSeg seg;                     // separable time dependence only: the knot interval the orbit's clock is in
};

// pull one queue position per calling lane (called from divergent code: aggregate over the
// lanes that are here together, one atomic per group)
__device__ __forceinline__ int64_t queue_pop(unsigned long long *queue)
{
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(queue, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    return (int64_t)(base + g.thread_rank());
}

// apply every kick scheduled on node s.j (kicks.py:160-172); returns true if the segment restarts
__device__ __forceinline__ bool apply_kicks(const IntegArgs &A, Lane &s)
{
    bool kicked = false;
    while (s.e < s.e1 && A.ev_step[s.e] == s.j) {
        s.vx += A.ev_dv[3 * s.e + 0];
        s.vy += A.ev_dv[3 * s.e + 1];
        s.vz += A.ev_dv[3 * s.e + 2];
        // gala restarts with dt = t[1] - t[0] of the new segment, both converted to Myr
        const double tk = A.ev_tk[s.e];
        const double tn = (s.append && s.j + 1 == s.L) ? s.t2 : tk + s.hgrid;
        s.dt = seg_dt_myr(tk, tn, s.smyr);
        kicked = true;
        s.e++;
    }
    return kicked;
}

__device__ __forceinline__ void write_final(const IntegArgs &A, const Lane &s)
{
    const int64_t n = A.n;
    A.w_final[s.oid] = s.x;
    A.w_final[n + s.oid] = s.y;
    A.w_final[2 * n + s.oid] = s.z;
    A.w_final[3 * n + s.oid] = s.vx;
    A.w_final[4 * n + s.oid] = s.vy;
    A.w_final[5 * n + s.oid] = s.vz;
}

template <int KIND, bool STORE>
__global__ void __launch_bounds__(STORE ? LF_BLOCK_STORE : LF_BLOCK) __maxnreg__(STORE ? (LF_BLOCK_STORE == 512 ? 128 : CW_LF_STORE_REGS) : (KIND == POT_GENERAL || KindTraits<KIND>::TS) ? CW_LF_GEN_REGS : CW_LF_REGS) leapfrog_kernel(const __grid_constant__ PotParams P,
                                                            const __grid_constant__ IntegArgs A)
{
    const unsigned FULL = 0xffffffffu;
    __shared__ double2 ltab_mem[LOG_TABLE_N];
    extern __shared__ __align__(16) unsigned char traj_smem[];
    const LogTab ltab = log_table_init(ltab_mem);
    __shared__ uint32_t tmem_slot;
    uint32_t tmem_base = 0;
    TrajWriter tw;
    if (STORE) {
        tw.init(traj_smem, A);
#if CW_TRAJ_TMEM
        // tensor memory (16 warps x 96 columns, or 8 x 192): second staging level of the trajectory writer
        tmem_base = traj_tmem_alloc(&tmem_slot);
        tw.enable_tmem(tmem_base);
#endif
        if (TRAJ_TSTAGE) {
            // tensor memory as the staging area of the upper half of the CTA's warps (cw_traj.cuh)
            tmem_base = traj_tmem_alloc(&tmem_slot);
            if ((int)threadIdx.x >= TRAJ_SMEM_LANES) tw.enable_tmem_stage(tmem_base);
        }
    }
    int64_t col_base = 0;          // STORE: column of sample 0 of the current orbit
    Lane s;
    s.x = 8.0; s.y = 0.0; s.z = 0.0; s.vx = s.vy = s.vz = 0.0;
    s.vhx = s.vhy = s.vhz = 0.0; s.Fxy = s.Fy = s.Fz = 0.0;
    s.dt = 0.0; s.hgrid = 0.0; s.t2 = 0.0; s.smyr = 1.0; s.tcur = 0.0;
    s.e = s.e1 = 0; s.oid = -1; s.j = 0; s.next = 0; s.L = 0; s.append = false;
    amp_seg_reset(s.seg);
    bool active = false, exhausted = false, staged = false;
    unsigned long long steps_done = 0;
    // STORE: the warp clock.  W counts the steps this warp has executed; an orbit whose first sample lives
    // in column c starts only when W = c (mod 16), so (column of the node a lane stands on) mod 16 is the
    // same for every active lane of the warp at every step: windows fill and flush in lock-step and the
    // lower / upper halves of the 128-byte lines alternate warp-wide.  A popped orbit waits at most 15 steps.
    unsigned W = 0;        // wraps modulo 2^32, a multiple of 16
    int delay = 0;
    bool pending = false;
    int64_t pend_o = -1;
    // STORE: the warp takes queue entries in private blocks of 32 CONSECUTIVE positions (one atomic per
    // block) and hands them to its lanes as they fall idle, so a warp always writes 32 neighbouring
    // trajectories even when its lanes finish at different steps.  (Per-lane pops interleave the warps:
    // after the first round every lane of a warp would hold an orbit from somewhere else in the queue,
    // and streams >= 32 KB apart cost a third of the HBM write throughput -- measured.)
    int64_t wq_next = 0, wq_end = 0;
    // De-phasing of the store bursts (STORE only).  With equal-length orbits every warp of the GPU
    // completes its staging windows on the same step, so the whole chip alternates between an FP64-only
    // phase (HBM idle) and a store burst (FP64 idle): measured 8.3 ms of arithmetic + 6.3 ms of stores
    // per 10^9 samples, strictly serialised.  Each warp therefore times its own first window period and
    // then waits once for (warp id mod 16) / 16 of it: from then on a sixteenth of the warps (one of the 16 on
    // every SM) is flushing at any moment while the others compute, and since all warps advance at the same
    // rate the offsets stay.  (16 offsets measured 2-3 % faster than 8 on config 3.)
    const unsigned gw = STORE ? (blockIdx.x * (LF_BLOCK_STORE / 32) + (threadIdx.x >> 5)) : 0u;
    long long t_first = STORE ? clock64() : 0;
    int dephase = STORE ? 2 : 0;       // 2: first flush not seen yet, 1: second flush pending, 0: done
    // Every force evaluation of the kernel.  The factors of the PREVIOUS node die here, so their register pairs serve as
    // the seed donors of the gradient (LogTab::d0, d1: two moves fewer per step).  Only low words are read and they are
    // a function of the orbit's own history -- an orbit starts with both factors zero -- so results stay independent
    // of lane, batch and kernel (final-state and trajectory runs evaluate the same sequence).
    auto eval_force = [&](double t) {
        LogTab lt = ltab;
        lt.d0 = s.Fxy;
        lt.d1 = s.Fz;
        force3t_seg<KIND>(P, lt, s.seg, t, s.x, s.y, s.z, s.Fxy, s.Fy, s.Fz);
    };

    for (;;) {
        // ---- phase A: lanes standing on a sync node (kick node / last node).  When STORE, the sample of
        // the node every active lane stands on has not been staged yet (it must carry the post-kick
        // velocity): stage it here, sync node or not.
        bool put_a = false;         // TRAJ_TSTAGE: the sample of phase A is staged by the whole warp together, below
        if (active && (s.j == s.next || (STORE && !staged))) {
            const bool at_sync = (s.j == s.next);
            const bool kicked = at_sync ? apply_kicks(A, s) : false;
            if (STORE) {
                if (TRAJ_TSTAGE) put_a = true;
                else tw.put(col_base + s.j, s.x, s.y, s.z, s.vx, s.vy, s.vz, s.j == s.L);
                staged = true;
            }
            if (at_sync) {
                if (s.j == s.L) {
                    write_final(A, s);
                    active = false;
                } else {
                    if (kicked) { // c_init_velocity of the new segment; the force at this node is already known
                        const double nhdt = -0.5 * s.dt;
                        s.vhx = lf_kick<KIND>(s.Fxy, nhdt, s.x, s.vx);
                        s.vhy = lf_kick<KIND>(s.Fy, nhdt, s.y, s.vy);
                        s.vhz = lf_kick<KIND>(s.Fz, nhdt, s.z, s.vz);
                    }
                    s.next = (s.e < s.e1) ? A.ev_step[s.e] : s.L;
                }
            }
        }
        if (STORE && TRAJ_TSTAGE)   // (position and synchronised velocity of the node are untouched by the block above)
            tw.put_any(put_a, (int)(W & (TRAJ_S - 1)), col_base + s.j, s.x, s.y, s.z, s.vx, s.vy, s.vz, s.j == s.L);
        if (STORE) {
            const bool flushed = tw.flush_warp(A);   // windows completed on this node (incl. the last of an orbit)
            if (dephase && flushed) {
                const long long now = clock64();
                if (dephase == 2) {
                    t_first = now;                   // start of the first full window period
                    dephase = 1;
                } else {
                    const long long wait = (now - t_first) * (long long)(gw % (2 * TRAJ_S)) / (2 * TRAJ_S);
                    while (clock64() - now < wait) { }
                    dephase = 0;
                }
            }
        }
        // ---- phase B: idle lanes pull the next orbit from the queue
        bool put_b = false;         // TRAJ_TSTAGE: the first sample of a new orbit, staged by the whole warp after the loop
        do {
        if (STORE) {
            // warp-uniform hand-out from the warp's private block of queue positions
            for (;;) {
                const bool need = !active && !exhausted && !pending;
                const unsigned nm = __ballot_sync(FULL, need);
                if (nm == 0) break;
                if (wq_next >= wq_end) {
                    unsigned long long base = 0;
                    if ((threadIdx.x & 31) == 0) base = atomicAdd(A.queue, 32ull);
                    base = __shfl_sync(FULL, base, 0);
                    wq_next = min((int64_t)base, A.n);
                    wq_end = min((int64_t)base + 32, A.n);
                }
                const int avail = (int)(wq_end - wq_next);
                const int rank = __popc(nm & ((1u << (threadIdx.x & 31)) - 1u));
                if (need) {
                    if (avail == 0) {
                        exhausted = true;
                    } else if (rank < avail) {
                        const int64_t p = wq_next + rank;
                        pend_o = A.order ? (int64_t)A.order[p] : p;
                        pending = true;
                        delay = 0;
                        if (A.status[pend_o] == CW_ORBIT_OK && A.n_nodes[pend_o] > 1)
                            delay = (int)((unsigned)(A.traj_off[pend_o] - A.traj_base) - W) & (2 * TRAJ_S - 1);
                    }
                }
                wq_next += min(__popc(nm), avail);
            }
        }
        while (!active && !exhausted) {
            int64_t o;
            if (STORE) {
                if (!pending || delay > 0) break;   // nothing handed out yet / not this orbit's start slot yet
                o = pend_o;
                pending = false;
            } else {
                const int64_t p = queue_pop(A.queue);
                if (p >= A.n) { exhausted = true; break; }
                o = A.order ? (int64_t)A.order[p] : p;
            }
            CW_ASSERT(o >= 0 && o < A.n);
            const int64_t n = A.n;
            s.oid = o;
            s.x = A.w0[o]; s.y = A.w0[n + o]; s.z = A.w0[2 * n + o];
            s.vx = A.w0[3 * n + o]; s.vy = A.w0[4 * n + o]; s.vz = A.w0[5 * n + o];
            s.j = 0;
            s.L = A.n_nodes[o] - 1;
            if (A.status[o] != CW_ORBIT_OK || s.L < 0) { // rejected by the pre-pass: hand back w0
                write_final(A, s);
                continue;
            }
            s.hgrid = A.dt[o]; s.t2 = A.t2[o]; s.smyr = A.to_myr[o];
            s.append = (A.flags[o] & CW_GRID_APPEND_T2) != 0;
            s.tcur = A.t1[o];
            s.e = s.e1 = 0;
            if (A.ev_off) { s.e = A.ev_off[o] - A.ev_base; s.e1 = A.ev_off[o + 1] - A.ev_base; }
            {   // dt = t[1] - t[0] in Myr (leapfrog.pyx); node 1 is t2 itself when it was appended
                const double t0 = s.tcur;
                const double tn = (s.append && s.L == 1) ? s.t2 : t0 + s.hgrid;
                s.dt = seg_dt_myr(t0, tn, s.smyr);
            }
            apply_kicks(A, s);          // events at (or before) the first node: kicks.py:155-157
            if (STORE) {
                col_base = A.traj_off[o] - A.traj_base;
                CW_ASSERT(col_base >= 0 && col_base + s.L < A.traj_total && A.traj_off[o + 1] - A.traj_off[o] > s.L);
                if (s.L == 0) traj_store_direct(A, col_base, s.x, s.y, s.z, s.vx, s.vy, s.vz);
                else if (TRAJ_TSTAGE) put_b = true;
                else tw.put(col_base, s.x, s.y, s.z, s.vx, s.vy, s.vz, false);
                staged = true;
            }
            if (s.L == 0) { write_final(A, s); continue; }
            s.Fxy = 0.0; s.Fz = 0.0;        // (seed donors of the first gradient: not the previous orbit's factors)
            eval_force(s.tcur * s.smyr);
            const double nhdt = -0.5 * s.dt;
            s.vhx = lf_kick<KIND>(s.Fxy, nhdt, s.x, s.vx);
            s.vhy = lf_kick<KIND>(s.Fy, nhdt, s.y, s.vy);
            s.vhz = lf_kick<KIND>(s.Fz, nhdt, s.z, s.vz);
            s.next = (s.e < s.e1) ? A.ev_step[s.e] : s.L;
            active = true;
        }
        } while (STORE && __any_sync(FULL, !active && !pending && !exhausted));   // rejected orbits: hand out again
        if (STORE && TRAJ_TSTAGE)   // (an orbit starts when the warp clock stands on its first column's phase)
            tw.put_any(put_b, (int)(W & (TRAJ_S - 1)), col_base, s.x, s.y, s.z, s.vx, s.vy, s.vz, false);
        if (STORE) tw.flush_warp(A);   // a first sample can complete a window (column = 7 mod 8)
        if (!__any_sync(FULL, active || (STORE && pending))) break;

        // ---- phase C: a warp-uniform run of steps up to the nearest sync node of any lane
        int r = active ? (s.next - s.j) : ((STORE && pending) ? delay : INT_MAX);
        if (STORE && active) {
            // ... and not past the node that completes this lane's staging window (column = 7 mod 8):
            // inside the run samples are staged without any vote, the completing one at the next phase A
            const int c7 = (int)((col_base + s.j) & (TRAJ_S - 1));
            r = min(r, ((TRAJ_S - 2 - c7) & (TRAJ_S - 1)) + 1);
        }
        const int m = __reduce_min_sync(FULL, r);
        // One formula for every step of both modes (so a final-state run equals the last sample of a
        // store_all run bit for bit):  x += vh dt;  F = force(x);  v = vh - grad dt/2;  vh -= grad dt,
        // with grad = (Fxy x, Fxy y, Fz z) folded into the FMAs: vh = fma(Fxy * (-dt), x, vh).
        const double dt = s.dt, ndt = -s.dt, nhdt = -0.5 * s.dt;
        constexpr bool TRI = (KindTraits<KIND>::BASE == POT_GENERIC);    // only the generic composite can have Fy != Fx
        constexpr bool TS = KindTraits<KIND>::TS;      // Phi = f(t) Phi_0: the factors are scaled by f(t_node)
        constexpr bool GEN = KindTraits<KIND>::VEC;    // (Fxy, Fy, Fz) is the gradient itself, evaluated at the node's time
        // time [Myr] of node J of this lane's orbit (POT_GENERAL only: the gradient of step j-1 -> j is taken at t_j)
        auto tnode = [&](int J) -> double {
            if (!GEN && !TS) return 0.0;
            return ((s.append && J == s.L) ? s.t2 : fma((double)J, s.hgrid, s.tcur)) * s.smyr;
        };
        if (GEN) {
            for (int q = 1; q < m; q++) {
                s.x = fma(s.vhx, dt, s.x);
                s.y = fma(s.vhy, dt, s.y);
                s.z = fma(s.vhz, dt, s.z);
                eval_force(tnode(s.j + q));
                if (STORE) {
                    s.vx = fma(s.Fxy, nhdt, s.vhx);
                    s.vy = fma(s.Fy, nhdt, s.vhy);
                    s.vz = fma(s.Fz, nhdt, s.vhz);
                }
                s.vhx = fma(s.Fxy, ndt, s.vhx);
                s.vhy = fma(s.Fy, ndt, s.vhy);
                s.vhz = fma(s.Fz, ndt, s.vhz);
                if (STORE && TRAJ_TSTAGE) tw.stage_any(active, (int)((W + (unsigned)q) & (TRAJ_S - 1)), col_base + s.j + q, s.x, s.y, s.z, s.vx, s.vy, s.vz);
                else if (STORE && active) tw.stage(col_base + s.j + q, s.x, s.y, s.z, s.vx, s.vy, s.vz);
            }
        } else if (!STORE) {
#pragma unroll 4
            for (int q = 1; q < m; q++) {   // c_leapfrog_step without the synchronised velocity
                s.x = fma(s.vhx, dt, s.x);
                s.y = fma(s.vhy, dt, s.y);
                s.z = fma(s.vhz, dt, s.z);
                eval_force(tnode(s.j + q));
                const double a = s.Fxy * ndt, b = s.Fz * ndt;
                const double ay = TRI ? s.Fy * ndt : a;
                s.vhx = fma(a, s.x, s.vhx);
                s.vhy = fma(ay, s.y, s.vhy);
                s.vhz = fma(b, s.z, s.vhz);
            }
        } else {
#ifdef CW_STORE_UNROLL
#define CW_PRAGMA_(x) _Pragma(#x)
#define CW_PRAGMA(x) CW_PRAGMA_(x)
            CW_PRAGMA(unroll CW_STORE_UNROLL)
#endif
            for (int q = 1; q < m; q++) {   // full steps, every sample staged
                s.x = fma(s.vhx, dt, s.x);
                s.y = fma(s.vhy, dt, s.y);
                s.z = fma(s.vhz, dt, s.z);
                eval_force(tnode(s.j + q));
                const double ah = s.Fxy * nhdt, bh = s.Fz * nhdt;
                const double ahy = TRI ? s.Fy * nhdt : ah;
                s.vx = fma(ah, s.x, s.vhx);
                s.vy = fma(ahy, s.y, s.vhy);
                s.vz = fma(bh, s.z, s.vhz);
                const double a = s.Fxy * ndt, b = s.Fz * ndt;
                const double ay = TRI ? s.Fy * ndt : a;
                s.vhx = fma(a, s.x, s.vhx);
                s.vhy = fma(ay, s.y, s.vhy);
                s.vhz = fma(b, s.z, s.vhz);
                if (TRAJ_TSTAGE) tw.stage_any(active, (int)((W + (unsigned)q) & (TRAJ_S - 1)), col_base + s.j + q, s.x, s.y, s.z, s.vx, s.vy, s.vz);
                else if (active) tw.stage(col_base + s.j + q, s.x, s.y, s.z, s.vx, s.vy, s.vz);
            }
        }
        if (GEN) {
            s.x = fma(s.vhx, dt, s.x);
            s.y = fma(s.vhy, dt, s.y);
            s.z = fma(s.vhz, dt, s.z);
            eval_force(tnode(s.j + m));
            s.vx = fma(s.Fxy, nhdt, s.vhx);
            s.vy = fma(s.Fy, nhdt, s.vhy);
            s.vz = fma(s.Fz, nhdt, s.vhz);
            s.vhx = fma(s.Fxy, ndt, s.vhx);
            s.vhy = fma(s.Fy, ndt, s.vhy);
            s.vhz = fma(s.Fz, ndt, s.vhz);
        } else {   // full c_leapfrog_step: v = vh - g dt/2 ; vh = vh - g dt
            s.x = fma(s.vhx, dt, s.x);
            s.y = fma(s.vhy, dt, s.y);
            s.z = fma(s.vhz, dt, s.z);
            eval_force(tnode(s.j + m));
            const double ah = s.Fxy * nhdt, bh = s.Fz * nhdt;
            const double ahy = TRI ? s.Fy * nhdt : ah;
            s.vx = fma(ah, s.x, s.vhx);
            s.vy = fma(ahy, s.y, s.vhy);
            s.vz = fma(bh, s.z, s.vhz);
            const double a = s.Fxy * ndt, b = s.Fz * ndt;
            const double ay = TRI ? s.Fy * ndt : a;
            s.vhx = fma(a, s.x, s.vhx);
            s.vhy = fma(ay, s.y, s.vhy);
            s.vhz = fma(b, s.z, s.vhz);
        }
        staged = false;
        if (active) {
            s.j += m;
            steps_done += (unsigned long long)m;
        }
        if (STORE) {
            W += (unsigned)m;
            if (pending) delay -= m;
        }
    }
#if CW_TRAJ_TMEM
    if (STORE) traj_tmem_free(tmem_base);
#endif
    if (STORE && TRAJ_TSTAGE) traj_tmem_free(tmem_base);
    if (A.stats) {
        // orbit-steps actually integrated (the bench's unit of work), one atomic per warp
        for (int off = 16; off > 0; off >>= 1) steps_done += __shfl_down_sync(FULL, steps_done, off);
        if ((threadIdx.x & 31) == 0 && steps_done) atomicAdd(A.stats + 3, steps_done);
    }
}

template <int KIND, bool STORE>
static cudaError_t launch_lf(const PotParams &P, const IntegArgs &A, const LaunchCfg &cfg)
{
    int per_sm = 0;
    int block = STORE ? LF_BLOCK_STORE : LF_BLOCK;
    // Batches too small to give every SM two full CTAs are cut into smaller CTAs (down to 64 threads), so the
    // orbits spread evenly over the 148 SMs instead of filling a few of them (results do not depend on it).
    if (!STORE)
        while (block > 64 && (A.n + block - 1) / block < 2 * (int64_t)cfg.sm_count) block >>= 1;
    const size_t dyn = STORE ? (size_t)TRAJ_SMEM_LANES * TrajWriter::LANE_BYTES : 0;
    cudaError_t err = cudaSuccess;
    if (STORE) {
        err = cudaFuncSetAttribute(leapfrog_kernel<KIND, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (err != cudaSuccess) return err;
    }
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, leapfrog_kernel<KIND, STORE>, block, dyn);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("CW_LF_MAX_CTAS")) {      // experiment hook: fewer resident CTAs per SM
        const int v = atoi(e);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    int64_t grid = (int64_t)cfg.sm_count * per_sm;
    const int64_t need = (A.n + block - 1) / block;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (getenv("CW_DEBUG"))
        fprintf(stderr, "[cw] leapfrog_kernel<%d,%d>: %d CTAs/SM x %d SMs, grid %lld, block %d, dyn smem %zu\n", KIND,
                (int)STORE, per_sm, cfg.sm_count, (long long)grid, block, dyn);
    leapfrog_kernel<KIND, STORE><<<(unsigned)grid, block, dyn, cfg.stream>>>(P, A);
    return cudaGetLastError();
}

template <int KIND>
static int64_t lf_capacity(int sm_count)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, leapfrog_kernel<KIND, false>, LF_BLOCK, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (int64_t)sm_count * per_sm * LF_BLOCK;
}

int64_t leapfrog_lane_capacity(const PotParams &P, int sm_count)
{
    switch (P.kind) {
    case POT_MW_V1: return lf_capacity<POT_MW_V1>(sm_count);
    case POT_MW_V2: return lf_capacity<POT_MW_V2>(sm_count);
    case POT_GENERAL: return lf_capacity<POT_GENERAL>(sm_count);
    case POT_ROTATED: return lf_capacity<POT_ROTATED>(sm_count);
    case POT_LM10: return lf_capacity<POT_LM10>(sm_count);
    case POT_MW_V2_T: return lf_capacity<POT_MW_V2_T>(sm_count);
    case POT_GENERIC_T: return lf_capacity<POT_GENERIC_T>(sm_count);
    default: return lf_capacity<POT_GENERIC>(sm_count);
    }
}

cudaError_t launch_leapfrog(const PotParams &P, const IntegArgs &A, bool store_all, const LaunchCfg &cfg)
{
    if (A.n <= 0) return cudaSuccess;
    switch (P.kind) {
    case POT_MW_V1:
        return store_all ? launch_lf<POT_MW_V1, true>(P, A, cfg) : launch_lf<POT_MW_V1, false>(P, A, cfg);
    case POT_MW_V2:
        return store_all ? launch_lf<POT_MW_V2, true>(P, A, cfg) : launch_lf<POT_MW_V2, false>(P, A, cfg);
    case POT_GENERAL:
        return store_all ? launch_lf<POT_GENERAL, true>(P, A, cfg) : launch_lf<POT_GENERAL, false>(P, A, cfg);
    case POT_ROTATED:
        return store_all ? launch_lf<POT_ROTATED, true>(P, A, cfg) : launch_lf<POT_ROTATED, false>(P, A, cfg);
    case POT_LM10:
        return store_all ? launch_lf<POT_LM10, true>(P, A, cfg) : launch_lf<POT_LM10, false>(P, A, cfg);
    case POT_MW_V2_T:
        return store_all ? launch_lf<POT_MW_V2_T, true>(P, A, cfg) : launch_lf<POT_MW_V2_T, false>(P, A, cfg);
    case POT_GENERIC_T:
        return store_all ? launch_lf<POT_GENERIC_T, true>(P, A, cfg) : launch_lf<POT_GENERIC_T, false>(P, A, cfg);
    default:
        return store_all ? launch_lf<POT_GENERIC, true>(P, A, cfg) : launch_lf<POT_GENERIC, false>(P, A, cfg);
    }
}

} // namespace cw
