// cw_traj.cuh -- full-trajectory output (store_all=True) in the reference's orbits/{pos,vel} layout
// (cogsworth pop.py:1853-1876): row c of traj_pos / traj_vel is one coordinate of ALL orbits, sample j
// of orbit i sits in column traj_off[i] + j.
//
// One lane integrates one orbit, so consecutive samples of a lane are consecutive in memory but the
// 32 lanes of a warp write 32 far-apart columns: a plain st.global.f64 per sample would cost a whole
// 32-byte L2 sector transaction for 8 useful bytes.  Instead every lane stages TRAJ_S (8) consecutive
// samples of its six coordinates in its own shared-memory rows.  Windows are aligned to ABSOLUTE
// columns (slot = column & 7), so a complete window is six 64-byte, 64-byte-aligned runs whatever the
// orbit's offset.  At the two warp-convergent points of the integrator loop the warp votes which lanes
// completed a window and writes them cooperatively: 8 adjacent lanes move one 64-byte run (LDS.64 ->
// STG.64), so every store instruction writes four full runs = eight full sectors, no partial-sector
// traffic.  The ragged first / last window of an orbit goes out through the same path, predicated.
//
// (Round-1 note: the first version issued per-lane cp.async.bulk (TMA) stores.  UBLKCP is a
// uniform-datapath instruction, so ptxas serialised the 32 lanes with an R2UR/PLOP3/BRA loop -- 466
// warp instructions per orbit-step, 2.2 TB/s.  See profiles/r1_ncu_store_all_tma_per_lane_summary.md.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "cw_kernels.cuh"

namespace cw {

constexpr int TRAJ_S = 8;                                    // samples per window
constexpr int TRAJ_ROWS = 6;                                 // x y z vx vy vz
constexpr int TRAJ_DATA_BYTES = TRAJ_ROWS * TRAJ_S * 8;      // 384: rows of one lane, [row][slot]
// + 8-byte header {column << 4 | count}.  392 bytes = 98 words = 2 (mod 32): the 8-byte staging stores
// of the 32 lanes fall into 32 distinct bank pairs (conflict-free), which a 16-byte-aligned stride
// cannot give -- so the flush moves 8-byte words (LDS.64 -> STG.64), 8 adjacent lanes per 64-byte run.
constexpr int TRAJ_LANE_BYTES = TRAJ_DATA_BYTES + 8;

struct TrajWriter {
    unsigned char *warp_base; // the 32 lane regions of this warp
    double *buf;              // this lane's rows: buf[row * TRAJ_S + slot]
    double *dst_pos, *dst_vel; // flush role of this lane: row 0 of pos / vel, offset by its slot (lane & 7)
    int64_t row_stride;        // columns per row
    int64_t col0;
    int n;
    bool want;                // my window is complete and waits for the next cooperative flush
    bool ok8;                 // destination rows are 8-byte aligned (always) -- kept for symmetry

    __device__ __forceinline__ void init(unsigned char *smem, const IntegArgs &A)
    {
        const int lane = threadIdx.x & 31;
        warp_base = smem + (size_t)(threadIdx.x - lane) * TRAJ_LANE_BYTES;
        buf = reinterpret_cast<double *>(warp_base + (size_t)lane * TRAJ_LANE_BYTES);
        // flush role: lanes 8q..8q+7 move the 64-byte run of source lane 4g + q (slot = lane & 7), one
        // row per instruction -- every store instruction writes four full runs
        dst_pos = A.traj_pos + (lane & 7);
        dst_vel = A.traj_vel + (lane & 7);
        row_stride = A.traj_total;
        col0 = 0;
        n = 0;
        want = false;
        ok8 = true;
    }

    // stage a sample that is known NOT to complete the window (the integrator bounds its runs so)
    __device__ __forceinline__ void stage(int64_t col, double x, double y, double z, double vx, double vy, double vz)
    {
        const int slot = (int)(col & (TRAJ_S - 1));
        if (n == 0) col0 = col;
        buf[0 * TRAJ_S + slot] = x;
        buf[1 * TRAJ_S + slot] = y;
        buf[2 * TRAJ_S + slot] = z;
        buf[3 * TRAJ_S + slot] = vx;
        buf[4 * TRAJ_S + slot] = vy;
        buf[5 * TRAJ_S + slot] = vz;
        n++;
    }

    // stage one sample for absolute column `col`; `last` = final sample of the orbit.
    // At most one window may complete between two flush_warp() calls.
    __device__ __forceinline__ void put(int64_t col, double x, double y, double z, double vx, double vy, double vz,
                                        bool last)
    {
        stage(col, x, y, z, vx, vy, vz);
        if ((col & (TRAJ_S - 1)) == TRAJ_S - 1 || last) {
            *reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(buf) + TRAJ_DATA_BYTES) =
                (long long)(col0 << 4) | (long long)n;
            want = true;
        }
    }

    // Warp-convergent: every lane of the warp must call this together.  Returns true if anything was
    // written (warp-uniform).
    __device__ __forceinline__ bool flush_warp(const IntegArgs &A)
    {
        unsigned m = __ballot_sync(0xffffffffu, want);
        if (m == 0) return false;
        __syncwarp();                       // the staged samples / headers of the other lanes are visible
        const int lane = threadIdx.x & 31;
        const int q = lane >> 3, slot = lane & 7;
        if (m == 0xffffffffu && __all_sync(0xffffffffu, n == TRAJ_S)) {
            // lockstep case (equal-length orbits): all 32 windows complete.  Fully unrolled, immediate
            // shared-memory offsets; per group of 4 source lanes ONE header load, then six independent
            // (LDS.64 -> STG.64) pairs, each writing four full 64-byte runs.
            const unsigned char *src = warp_base + q * TRAJ_LANE_BYTES + slot * 8;
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const unsigned char *sg = src + g * 4 * TRAJ_LANE_BYTES;
                const long long c0 = *reinterpret_cast<const long long *>(sg - slot * 8 + TRAJ_DATA_BYTES) >> 4;
                double *dp = dst_pos + c0, *dv = dst_vel + c0;
                const double w0 = *reinterpret_cast<const double *>(sg + 0 * TRAJ_S * 8);
                const double w1 = *reinterpret_cast<const double *>(sg + 1 * TRAJ_S * 8);
                const double w2 = *reinterpret_cast<const double *>(sg + 2 * TRAJ_S * 8);
                const double w3 = *reinterpret_cast<const double *>(sg + 3 * TRAJ_S * 8);
                const double w4 = *reinterpret_cast<const double *>(sg + 4 * TRAJ_S * 8);
                const double w5 = *reinterpret_cast<const double *>(sg + 5 * TRAJ_S * 8);
#ifndef CW_TRAJ_NOSTORE
                dp[0] = w0;
                dp[row_stride] = w1;
                dp[2 * row_stride] = w2;
                dv[0] = w3;
                dv[row_stride] = w4;
                dv[2 * row_stride] = w5;
#else
                if (c0 == -12345) dp[0] = w0 + w1 + w2 + w3 + w4 + w5;
#endif
            }
        } else {
            while (m) {
                // up to four flushing source lanes per trip, one per 8-lane group
                int l = -1;
                unsigned mm = m;
                for (int k = 0; k < 4; k++) {
                    const int cand = mm ? (__ffs(mm) - 1) : -1;
                    if (mm) mm &= mm - 1;
                    if (k == q) l = cand;
                }
                m = mm;
                if (l >= 0) {
                    const unsigned char *sl = warp_base + (size_t)l * TRAJ_LANE_BYTES;
                    const long long h = *reinterpret_cast<const long long *>(sl + TRAJ_DATA_BYTES);
                    const long long c0 = h >> 4;
                    const int cnt = (int)(h & 15);
                    const int s0 = (int)(c0 & (TRAJ_S - 1));
                    if (slot >= s0 && slot < s0 + cnt) {
                        const double *rowbuf = reinterpret_cast<const double *>(sl) + slot;
                        double *dp = dst_pos + (c0 - s0), *dv = dst_vel + (c0 - s0);
#pragma unroll
                        for (int r = 0; r < 3; r++) {
                            dp[r * row_stride] = rowbuf[r * TRAJ_S];
                            dv[r * row_stride] = rowbuf[(3 + r) * TRAJ_S];
                        }
                    }
                }
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
    A.traj_pos[col] = x;
    A.traj_pos[A.traj_total + col] = y;
    A.traj_pos[2 * A.traj_total + col] = z;
    A.traj_vel[col] = vx;
    A.traj_vel[A.traj_total + col] = vy;
    A.traj_vel[2 * A.traj_total + col] = vz;
}

} // namespace cw
