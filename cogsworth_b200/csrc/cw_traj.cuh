// cw_traj.cuh -- full-trajectory output (store_all=True) in the reference's orbits/{pos,vel} layout
// (cogsworth pop.py:1853-1876): row c of traj_pos / traj_vel is one coordinate of ALL orbits, sample j
// of orbit i sits in column traj_off[i] + j.
//
// One lane integrates one orbit, so consecutive samples of a lane are consecutive in memory but the
// 32 lanes of a warp write 32 far-apart columns: a plain st.global.f64 per sample would cost a whole
// 32-byte L2 sector transaction for 8 useful bytes.  Instead every lane stages TRAJ_S (16; 8 in round 1)
// consecutive samples of its six coordinates.  Windows are aligned to ABSOLUTE columns (slot = column &
// 15), so a complete window is six full, aligned 128-byte lines whatever the orbit's offset.  Where the
// windows live depends on the warp (TRAJ_TSTAGE, the shipped form: 512 lanes per SM):
//   * warps 0..7: the lane's own shared-memory rows; at the two warp-convergent points of the integrator loop
//     the warp votes which lanes completed a window and writes them cooperatively, 16 adjacent lanes per
//     128-byte run (LDS.64 -> STG.64), two full lines per store instruction;
//   * warps 8..15: tensor memory, written straight from registers (tcgen05.st, two columns per value of the
//     thread's own TMEM lane) and read back in the 16x256b shape, which hands four adjacent threads four
//     consecutive samples of one orbit: every store instruction writes eight full 32-byte sectors.
// The ragged first / last window of an orbit goes out through the same paths, predicated.
//
// Lock-step is kept by construction, not by luck: the integrator starts an orbit only on a step of its
// warp clock W with W = (first column of the orbit) mod 32, so every active lane of a warp always stands
// on the same column phase, whatever the queue order and however different the orbit lengths are.
//
// (Round-1 note: the first version issued per-lane cp.async.bulk (TMA) stores.  UBLKCP is a
// uniform-datapath instruction, so ptxas serialised the 32 lanes with an R2UR/PLOP3/BRA loop -- 466
// warp instructions per orbit-step, 2.2 TB/s.  See profiles/r1_ncu_store_all_tma_per_lane_summary.md.)
#pragma once
#include <cassert>
#include <cstdint>
#include <cuda_runtime.h>
#include "cw_kernels.cuh"

// CW_CHECKED=1 (build knob): device-side asserts on every index the trajectory writer forms -- the stand-in for
// compute-sanitizer memcheck, which is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.md)
#ifdef CW_CHECKED
#define CW_ASSERT(x) assert(x)
#else
#define CW_ASSERT(x) ((void)0)
#endif

namespace cw {

#ifndef CW_TRAJ_WIDE
#define CW_TRAJ_WIDE 0     // build knob: pair staging + 16-byte flush for 16-sample windows.  Measured at full size
                           // (config 3, 48 GB): 3.77 TB/s against 4.07 TB/s for the 8-byte scheme -- half the shared-
                           // memory instructions, but the kernel is bound by dependent-issue latency at 8 warps per SM
                           // and by the 128-byte-run write pattern (ceiling 4.5 TB/s), not by MIO; off.
#endif
#ifndef CW_TRAJ_S
#define CW_TRAJ_S 16       // build knob: 8 = round-1 geometry (64-byte runs, 512 lanes per SM), 16 = 128-byte runs, 256 lanes
#endif
constexpr int TRAJ_S = CW_TRAJ_S;                            // samples per window
static_assert(TRAJ_S == 8 || TRAJ_S == 16, "TRAJ_S must be 8 or 16");
constexpr int TRAJ_ROWS = 6;                                 // x y z vx vy vz
// One lane's rows [row][slot] + an 8-byte header {column << 5 | count}: 392 (776) bytes = 98 (194) words = 2 (mod 32): the 8-byte staging stores
// of the 32 lanes fall into 32 distinct bank pairs (conflict-free), which a 16-byte-aligned stride
// cannot give -- so the flush moves 8-byte words (LDS.64 -> STG.64), TRAJ_S adjacent lanes per run of TRAJ_S * 8 bytes.
constexpr int TRAJ_HDR_SHIFT = 5;                            // header = first column << 5 | samples in the window (<= 16)
// Lanes of the trajectory-writing leapfrog kernel per SM.  Everything staged per lane is 2 x TRAJ_S samples x 48 B (one
// window in shared memory, one parked in tensor memory): 512 lanes x 16 samples = 64-byte runs (round 1), 256 lanes x
// 32 samples = 128-byte runs that leave in pairs, 256 contiguous bytes per stream.  The pure store pattern reaches 4.8
// TB/s with 128-byte runs and 6.5 TB/s with 256-byte ones (cw_measure_scatter_write), and 8 warps per SM still sustain
// 1.33e11 orbit-steps/s of the leapfrog arithmetic (tools/gpu_occ_probe.py) = 6.4 TB/s of samples.
// CW_TRAJ_TSTAGE (16-sample windows only): 512 lanes per SM.  Shared memory holds the windows of 256 lanes (196 KB); the
// other 256 lanes stage theirs in TENSOR MEMORY, written straight from registers (tcgen05.st, two columns per value)
// and read back at the flush in the 16x256b shape, which hands four adjacent threads four consecutive samples of one
// orbit: every store instruction writes eight full 32-byte sectors.  16 warps per SM instead of 8: the arithmetic runs
// at the rate of the final-state kernel and a warp that is flushing leaves three others on its scheduler computing.
#ifndef CW_TRAJ_TSTAGE
#define CW_TRAJ_TSTAGE 1
#endif
constexpr bool TRAJ_TSTAGE = (CW_TRAJ_TSTAGE != 0) && (TRAJ_S == 16);
constexpr int TRAJ_SMEM_LANES = (TRAJ_S == 8) ? 512 : 256;   // lanes whose windows live in shared memory
constexpr int TRAJ_STORE_BLOCK = TRAJ_TSTAGE ? 512 : TRAJ_SMEM_LANES;

// ---- tensor memory as a second staging level ---------------------------------------------------------
// Shared memory caps the staging windows at 8 samples = 64-byte runs, and lock-step 64-byte appends to
// millions of streams reach only 2.6-4.0 TB/s of HBM (cw_measure_scatter_write): by the time the second
// half of a 128-byte line arrives, a whole sweep later, the first half has long been evicted from L2 and
// DRAM sees half-line writes.  The 256 KB of TMEM per SM are idle in this kernel, so a completed LOWER
// half-line window (columns 0..7 mod 16) is parked there (tcgen05.st, 96 columns per lane) instead of
// being written, and is written right after the UPPER half 8 steps later: the two halves reach L2 within
// a microsecond of each other, merge, and leave as full 128-byte lines.
// Parking is decided warp-wide (tcgen05.st is .sync.aligned) but released PER LANE: when one lane's orbit ends
// in the middle of a window, only that lane's parked half is written; the others stay in TMEM until their own
// upper window completes, because their shared-memory rows hold the samples of the window still filling.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
                 "%13, %14, %15, %16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                 "%14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}

// one FP64 value = two 32-bit columns of the calling thread's own TMEM lane
__device__ __forceinline__ void tmem_st_f64(uint32_t taddr, double d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 :: "r"(taddr), "r"((uint32_t)__double2loint(d)), "r"((uint32_t)__double2hiint(d)) : "memory");
}

// 16 TMEM lanes x 32 columns in the 16x256b shape: thread t receives, for k = 0..3, columns 8k + 2 (t % 4) + {0, 1} of
// lane t / 4 in v[4k], v[4k+1] and of lane t / 4 + 8 in v[4k+2], v[4k+3]
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                 "%14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}

constexpr int TRAJ_TMEM_COLS_PER_CTA = 512;                        // 16 (8) warps: 4 (2) per lane quadrant x 96 (192) = 384 -> 512

// one warp of the CTA allocates; returns the TMEM base address to every thread (0 = not allocated)
__device__ __forceinline__ uint32_t traj_tmem_alloc(uint32_t *slot)
{
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(slot)), "n"(TRAJ_TMEM_COLS_PER_CTA) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}

__device__ __forceinline__ void traj_tmem_free(uint32_t base)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(TRAJ_TMEM_COLS_PER_CTA)
                     : "memory");
}

template <int S>
struct TrajWriterT {
    static_assert(S == 8 || S == 16, "window of 8 or 16 samples");
    static constexpr int DATA_BYTES = TRAJ_ROWS * S * 8;       // rows of one lane, [row][slot]
    // + header.  392 / 776 B = 98 / 194 words = 2 (mod 32): conflict-free for the 8-byte staging stores.  WIDE (knob, off):
    // samples staged in PAIRS with 16-byte stores and flushed with 16-byte loads / stores, which wants 16-byte aligned
    // rows and a stride that is an odd number of 16-byte units: 784 B = 49 units.
    static constexpr bool WIDE = (S == 16) && (CW_TRAJ_WIDE != 0);
    static constexpr int LANE_BYTES = DATA_BYTES + (WIDE ? 16 : 8);
    static constexpr int RPI = 32 / S;                         // runs one 8-byte store instruction of a warp writes
    static constexpr int TMEM_COLS_PER_WARP = TRAJ_ROWS * S * 2;

    unsigned char *warp_base; // the 32 lane regions of this warp
    double *buf;              // this lane's rows: buf[row * S + slot]
    double *dst_pos, *dst_vel; // flush role of this lane: row 0 of pos / vel, offset by its slot (lane & 7)
    int64_t row_stride;        // columns per row
    int64_t col0;
    int n;
    bool want;                // my window is complete and waits for the next cooperative flush
    bool want_last;           // ... and it is the last window of the orbit
    bool parked;              // warp-uniform: lower half-line windows of some lanes sit in TMEM
    bool lane_parked;         // ... and this lane is one of them
    int64_t park_col0;        // first column of this lane's parked window
    uint32_t taddr;           // this warp's TMEM block (lane quadrant << 16 | first column)
    bool tmem_on;             // TMEM parking enabled
    bool tm;                  // warp-uniform: this warp stages its windows in tensor memory (TRAJ_TSTAGE)
    // WIDE: the sample of an even column waits here for its odd neighbour (one 16-byte store for both)
    double hx, hy, hz, hvx, hvy, hvz;
    bool held;
    bool wide_ok;             // destination rows are 16-byte aligned: the lock-step flush may use 16-byte stores
    double *pos0, *vel0;

    __device__ __forceinline__ void init(unsigned char *smem, const IntegArgs &A)
    {
        const int lane = threadIdx.x & 31;
        // (TRAJ_TSTAGE: the warps beyond TRAJ_SMEM_LANES have no rows here and never touch these two)
        warp_base = smem + (size_t)((threadIdx.x - lane) % TRAJ_SMEM_LANES) * LANE_BYTES;
        buf = reinterpret_cast<double *>(warp_base + (size_t)lane * LANE_BYTES);
        // flush role: lanes S q .. S q + S-1 move the run of source lane RPI g + q (slot = lane % S), one
        // row per instruction -- every store instruction writes RPI full runs
        dst_pos = A.traj_pos + (lane & (S - 1));
        dst_vel = A.traj_vel + (lane & (S - 1));
        row_stride = A.traj_total;
        pos0 = A.traj_pos;
        vel0 = A.traj_vel;
        wide_ok = WIDE && ((reinterpret_cast<uintptr_t>(A.traj_pos) | reinterpret_cast<uintptr_t>(A.traj_vel)) & 15) == 0 &&
                  (A.traj_total & 1) == 0;
        hx = hy = hz = hvx = hvy = hvz = 0.0;
        held = false;
        col0 = 0;
        n = 0;
        want = want_last = parked = lane_parked = false;
        park_col0 = 0;
        taddr = 0;
        tmem_on = false;
        tm = false;
    }

    // TRAJ_TSTAGE: warps TRAJ_SMEM_LANES / 32 .. of the CTA stage in tensor memory.  Warp w addresses TMEM lanes
    // 32 (w % 4) .. +31 and owns 192 columns there (6 rows x 16 samples x 2).
    __device__ __forceinline__ void enable_tmem_stage(uint32_t tmem_base)
    {
        const int warp = threadIdx.x >> 5;
        const int k = (warp - TRAJ_SMEM_LANES / 32) >> 2;
        CW_ASSERT(k >= 0 && (k + 1) * TMEM_COLS_PER_WARP <= TRAJ_TMEM_COLS_PER_CTA);
        taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(k * TMEM_COLS_PER_WARP);
        tm = true;
    }

    // The warp-collective forms (TRAJ_TSTAGE): EVERY lane of the warp calls them together, `live` tells whether this
    // lane has a sample; `uslot` is the warp clock's column phase (the same for every lane; live lanes' own column
    // agrees with it).  Lanes without a sample write into their own, idle, TMEM rows.
    __device__ __forceinline__ void stage_any(bool live, int uslot, int64_t col, double x, double y, double z, double vx,
                                              double vy, double vz)
    {
        if (!tm) {
            if (live) stage(col, x, y, z, vx, vy, vz);
            return;
        }
        if (live) {
            if (n == 0) col0 = col;
            CW_ASSERT(col >= 0 && col < row_stride && n < S && col - col0 == n && (int)(col & (S - 1)) == uslot);
            n++;
        }
        const uint32_t a = taddr + 2u * (uint32_t)uslot;
        tmem_st_f64(a + 0 * 2 * S, x);
        tmem_st_f64(a + 1 * 2 * S, y);
        tmem_st_f64(a + 2 * 2 * S, z);
        tmem_st_f64(a + 3 * 2 * S, vx);
        tmem_st_f64(a + 4 * 2 * S, vy);
        tmem_st_f64(a + 5 * 2 * S, vz);
    }

    __device__ __forceinline__ void put_any(bool live, int uslot, int64_t col, double x, double y, double z, double vx,
                                            double vy, double vz, bool last)
    {
        if (!tm) {
            if (live) put(col, x, y, z, vx, vy, vz, last);
            return;
        }
        stage_any(live, uslot, col, x, y, z, vx, vy, vz);
        if (live && ((col & (S - 1)) == S - 1 || last)) {
            want = true;
            want_last = last;
        }
    }

    // flush of a TMEM-staging warp: headers travel by shuffle, data through tcgen05.ld in the 16x256b shape
    __device__ __forceinline__ bool flush_warp_tm()
    {
        const unsigned FULL = 0xffffffffu;
        const unsigned m = __ballot_sync(FULL, want);
        if (m == 0) return false;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const int lane = threadIdx.x & 31;
        const long long hdr = want ? ((long long)(col0 << TRAJ_HDR_SHIFT) | (long long)n) : 0ll;
        const int sub = lane & 3;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int La = half * 16 + (lane >> 2), Lb = La + 8;        // the two source lanes this thread serves
            const long long ha = __shfl_sync(FULL, hdr, La), hb = __shfl_sync(FULL, hdr, Lb);
            const int cnta = (int)(ha & ((1 << TRAJ_HDR_SHIFT) - 1)), cntb = (int)(hb & ((1 << TRAJ_HDR_SHIFT) - 1));
            const long long c0a = ha >> TRAJ_HDR_SHIFT, c0b = hb >> TRAJ_HDR_SHIFT;
            const int s0a = (int)(c0a & (S - 1)), s0b = (int)(c0b & (S - 1));
            CW_ASSERT(cnta == 0 || (c0a >= 0 && s0a + cnta <= S && c0a + cnta <= row_stride));
            CW_ASSERT(cntb == 0 || (c0b >= 0 && s0b + cntb <= S && c0b + cntb <= row_stride));
#pragma unroll
            for (int r = 0; r < TRAJ_ROWS; r++) {
                uint32_t v[16];
                tmem_ld_16x256b_x4(taddr + ((uint32_t)(half * 16) << 16) + (uint32_t)(2 * S * r), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                double *rowp = (r < 3) ? pos0 + r * row_stride : vel0 + (r - 3) * row_stride;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int sl = 4 * k + sub;
                    if (sl >= s0a && sl < s0a + cnta) rowp[(c0a - s0a) + sl] = __hiloint2double((int)v[4 * k + 1], (int)v[4 * k]);
                    if (sl >= s0b && sl < s0b + cntb) rowp[(c0b - s0b) + sl] = __hiloint2double((int)v[4 * k + 3], (int)v[4 * k + 2]);
                }
            }
        }
        if (want) {
            want = false;
            n = 0;
        }
        return true;
    }

    // tmem_base: value returned by traj_tmem_alloc.  Warp w owns TMEM lanes 32 (w % 4) .. +31 (the only ones
    // it can address) and, of those, columns 96 (w / 4) .. +95.
    __device__ __forceinline__ void enable_tmem(uint32_t tmem_base)
    {
        const int warp = threadIdx.x >> 5;
        CW_ASSERT(((warp >> 2) + 1) * TMEM_COLS_PER_WARP <= TRAJ_TMEM_COLS_PER_CTA);
        taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TMEM_COLS_PER_WARP);
        tmem_on = true;
    }

    // stage a sample that is known NOT to complete the window (the integrator bounds its runs so)
    __device__ __forceinline__ void stage(int64_t col, double x, double y, double z, double vx, double vy, double vz)
    {
        const int slot = (int)(col & (S - 1));
        if (n == 0) col0 = col;
        CW_ASSERT(col >= 0 && col < row_stride && n < S && col - col0 == n && (col0 & (S - 1)) + n < S);
        n++;
        if (WIDE) {
            if ((slot & 1) == 0) {          // even column: wait for the neighbour (an even column never ends a window)
                hx = x; hy = y; hz = z; hvx = vx; hvy = vy; hvz = vz;
                held = true;
                return;
            }
            if (held) {
                double2 *b2 = reinterpret_cast<double2 *>(buf + slot - 1);
                b2[0 * S / 2] = make_double2(hx, x);
                b2[1 * S / 2] = make_double2(hy, y);
                b2[2 * S / 2] = make_double2(hz, z);
                b2[3 * S / 2] = make_double2(hvx, vx);
                b2[4 * S / 2] = make_double2(hvy, vy);
                b2[5 * S / 2] = make_double2(hvz, vz);
                held = false;
                return;
            }
        }
        buf[0 * S + slot] = x;
        buf[1 * S + slot] = y;
        buf[2 * S + slot] = z;
        buf[3 * S + slot] = vx;
        buf[4 * S + slot] = vy;
        buf[5 * S + slot] = vz;
    }

    // WIDE: the held sample of an even column goes out on its own (its orbit ends on it)
    __device__ __forceinline__ void release_held(int64_t col)
    {
        if (WIDE && held) {
            const int slot = (int)(col & (S - 1));
            buf[0 * S + slot] = hx;
            buf[1 * S + slot] = hy;
            buf[2 * S + slot] = hz;
            buf[3 * S + slot] = hvx;
            buf[4 * S + slot] = hvy;
            buf[5 * S + slot] = hvz;
            held = false;
        }
    }

    // stage one sample for absolute column `col`; `last` = final sample of the orbit.
    // At most one window may complete between two flush_warp() calls.
    __device__ __forceinline__ void put(int64_t col, double x, double y, double z, double vx, double vy, double vz,
                                        bool last)
    {
        stage(col, x, y, z, vx, vy, vz);
        if ((col & (S - 1)) == S - 1 || last) {
            release_held(col);
            *reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(buf) + DATA_BYTES) =
                (long long)(col0 << TRAJ_HDR_SHIFT) | (long long)n;
            want = true;
            want_last = last;
        }
    }

    // own rows -> TMEM (warp-convergent)
    __device__ __forceinline__ void park()
    {
#pragma unroll
        for (int r = 0; r < TRAJ_ROWS; r++) {
#pragma unroll
            for (int h = 0; h < S / 8; h++) {      // 8 samples = 16 words per tcgen05.st
                uint32_t v[16];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const double d = buf[r * S + 8 * h + k];
                    v[2 * k] = (uint32_t)__double2loint(d);
                    v[2 * k + 1] = (uint32_t)__double2hiint(d);
                }
                tmem_st16(taddr + 2 * S * r + 16 * h, v);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        park_col0 = col0;
        lane_parked = want;     // lanes that had nothing to flush parked garbage: never written back
        parked = true;
    }

    // TMEM -> own rows + header, for the lanes with `mine` set; the caller then flushes them like a fresh
    // window.  Warp-convergent (tcgen05.ld is .sync.aligned: every lane loads), but ONLY the selected lanes
    // touch their shared-memory rows: a lane that is not flushing may hold the staged samples of its
    // current window there.
    __device__ __forceinline__ void unpark_to_smem(bool mine)
    {
#pragma unroll
        for (int r = 0; r < TRAJ_ROWS; r++) {
#pragma unroll
            for (int h = 0; h < S / 8; h++) {
                uint32_t v[16];
                tmem_ld16(taddr + 2 * S * r + 16 * h, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (mine) {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        buf[r * S + 8 * h + k] = __hiloint2double((int)v[2 * k + 1], (int)v[2 * k]);
                }
            }
        }
        if (mine)
            *reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(buf) + DATA_BYTES) =
                (long long)(park_col0 << TRAJ_HDR_SHIFT) | (long long)S;
    }

    // the lock-step store loop: all 32 lanes hold a complete window + header in shared memory
    __device__ __forceinline__ void store_all_windows(int lane)
    {
        const int q = lane / S, slot = lane & (S - 1);
        const unsigned char *src = warp_base + q * LANE_BYTES + slot * 8;
#pragma unroll
        for (int g = 0; g < S; g++) {
            const unsigned char *sg = src + g * RPI * LANE_BYTES;
            const long long c0 = *reinterpret_cast<const long long *>(sg - slot * 8 + DATA_BYTES) >> TRAJ_HDR_SHIFT;
            CW_ASSERT(c0 >= 0 && (c0 & (S - 1)) == 0 && c0 + S <= row_stride);
            double *dp = dst_pos + c0, *dv = dst_vel + c0;
            const double w0 = *reinterpret_cast<const double *>(sg + 0 * S * 8);
            const double w1 = *reinterpret_cast<const double *>(sg + 1 * S * 8);
            const double w2 = *reinterpret_cast<const double *>(sg + 2 * S * 8);
            const double w3 = *reinterpret_cast<const double *>(sg + 3 * S * 8);
            const double w4 = *reinterpret_cast<const double *>(sg + 4 * S * 8);
            const double w5 = *reinterpret_cast<const double *>(sg + 5 * S * 8);
            dp[0] = w0;
            dp[row_stride] = w1;
            dp[2 * row_stride] = w2;
            dv[0] = w3;
            dv[row_stride] = w4;
            dv[2 * row_stride] = w5;
        }
    }

    // WIDE form of the lock-step store loop: 8 adjacent lanes move one 128-byte run 16 bytes at a time (LDS.128 ->
    // STG.128), four runs per store instruction, 8 trips
    __device__ __forceinline__ void store_all_windows_wide(int lane)
    {
        const int q = lane >> 3, u = lane & 7;
        const unsigned char *src = warp_base + q * LANE_BYTES + u * 16;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const unsigned char *sg = src + g * 4 * LANE_BYTES;
            const long long c0 = *reinterpret_cast<const long long *>(sg - u * 16 + DATA_BYTES) >> TRAJ_HDR_SHIFT;
            CW_ASSERT(c0 >= 0 && (c0 & (S - 1)) == 0 && c0 + S <= row_stride);
            double *dp = pos0 + c0 + 2 * u, *dv = vel0 + c0 + 2 * u;
            const double2 w0 = *reinterpret_cast<const double2 *>(sg + 0 * S * 8);
            const double2 w1 = *reinterpret_cast<const double2 *>(sg + 1 * S * 8);
            const double2 w2 = *reinterpret_cast<const double2 *>(sg + 2 * S * 8);
            const double2 w3 = *reinterpret_cast<const double2 *>(sg + 3 * S * 8);
            const double2 w4 = *reinterpret_cast<const double2 *>(sg + 4 * S * 8);
            const double2 w5 = *reinterpret_cast<const double2 *>(sg + 5 * S * 8);
            *reinterpret_cast<double2 *>(dp) = w0;
            *reinterpret_cast<double2 *>(dp + row_stride) = w1;
            *reinterpret_cast<double2 *>(dp + 2 * row_stride) = w2;
            *reinterpret_cast<double2 *>(dv) = w3;
            *reinterpret_cast<double2 *>(dv + row_stride) = w4;
            *reinterpret_cast<double2 *>(dv + 2 * row_stride) = w5;
        }
    }

    // The same sweep for ANY subset of lanes and ragged windows (first / last window of an orbit, lanes
    // that are idle or waiting for their start slot): `mask` = source lanes that flush, each window's
    // extent comes from its header.  Eight trips whatever the mask, four 64-byte runs per store instruction.
    __device__ __forceinline__ void store_windows_masked(int lane, unsigned mask)
    {
        const int q = lane / S, slot = lane & (S - 1);
#pragma unroll
        for (int g = 0; g < S; g++) {
            const int l = RPI * g + q;
            if ((mask >> l) & 1u) {
                const unsigned char *sl = warp_base + (size_t)l * LANE_BYTES;
                const long long h = *reinterpret_cast<const long long *>(sl + DATA_BYTES);
                const long long c0 = h >> TRAJ_HDR_SHIFT;
                const int cnt = (int)(h & ((1 << TRAJ_HDR_SHIFT) - 1));
                const int s0 = (int)(c0 & (S - 1));
                CW_ASSERT(c0 >= 0 && cnt >= 1 && s0 + cnt <= S && c0 + cnt <= row_stride);
                if (slot >= s0 && slot < s0 + cnt) {
                    const double *rowbuf = reinterpret_cast<const double *>(sl) + slot;
                    double *dp = dst_pos + (c0 - s0), *dv = dst_vel + (c0 - s0);
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        dp[r * row_stride] = rowbuf[r * S];
                        dv[r * row_stride] = rowbuf[(3 + r) * S];
                    }
                }
            }
        }
    }

    // Warp-convergent: every lane of the warp must call this together.  Returns true if any window was
    // completed (warp-uniform).
    __device__ __forceinline__ bool flush_warp(const IntegArgs &A)
    {
        if (TRAJ_TSTAGE && S == 16 && tm) return flush_warp_tm();
        const unsigned FULL = 0xffffffffu;
        const unsigned m = __ballot_sync(FULL, want);
        if (m == 0) return false;
        __syncwarp();                       // the staged samples / headers of the other lanes are visible
        const int lane = threadIdx.x & 31;
        // a complete window on the lower half of a 128-byte line, not the last of its orbit, can wait in TMEM
        const bool parkable = want && n == S && (col0 & (2 * S - 1)) == 0 && !want_last;
        if (tmem_on && !parked && __ballot_sync(FULL, parkable) == m) {
            park();                         // every flushing lane parks; nothing is written
#ifdef CW_TRAJ_DEBUG
            if (lane == 0) atomicAdd(A.stats + 0, 1ull);
#endif
        } else {
            const bool lockstep = (m == FULL) && __all_sync(FULL, n == S);
#ifdef CW_TRAJ_DEBUG
            if (lane == 0) atomicAdd(A.stats + (lockstep ? 2 : 1), lockstep ? 1ull : (unsigned long long)(__popc(m) + 1000000ull));
#endif
            if (lockstep && wide_ok) store_all_windows_wide(lane);
            else if (lockstep) store_all_windows(lane);
            else store_windows_masked(lane, m);
            if (parked) {
                // The parked lower halves of the lanes that just flushed follow immediately, so both halves of
                // every line meet in L2.  Lanes that are NOT flushing now (their upper window is still filling:
                // another lane's orbit ended mid-window) keep theirs in TMEM -- their shared-memory rows hold
                // live samples -- and write it when their own window completes.
                const unsigned pm = __ballot_sync(FULL, lane_parked);
                const unsigned um = pm & m;
                if (um) {
                    const bool mine = (um >> lane) & 1u;
                    __syncwarp();           // the rows just written out are free again
                    unpark_to_smem(mine);
                    __syncwarp();
                    if (um == FULL && wide_ok) store_all_windows_wide(lane);
                    else if (um == FULL) store_all_windows(lane);
                    else store_windows_masked(lane, um);
                    if (mine) lane_parked = false;
                }
                parked = (pm & ~um) != 0u;
            }
        }
        __syncwarp();                       // rows may be overwritten by their owners from here on
        if (want) {
            want = false;
            n = 0;
        }
        return true;
    }
};

// unstaged store of a single sample (degenerate one-node orbits)
__device__ __forceinline__ void traj_store_direct(const IntegArgs &A, int64_t col, double x, double y, double z,
                                                  double vx, double vy, double vz)
{
    CW_ASSERT(col >= 0 && col < A.traj_total);
    A.traj_pos[col] = x;
    A.traj_pos[A.traj_total + col] = y;
    A.traj_pos[2 * A.traj_total + col] = z;
    A.traj_vel[col] = vx;
    A.traj_vel[A.traj_total + col] = vy;
    A.traj_vel[2 * A.traj_total + col] = vz;
}

// the leapfrog kernels write TRAJ_S-sample windows (build knob); DOPRI853 (compute-bound: 45 FLOP per stored byte) keeps
// the small windows so that two of its 128-thread CTAs fit an SM
using TrajWriter = TrajWriterT<TRAJ_S>;
using TrajWriterDop = TrajWriterT<8>;

} // namespace cw
