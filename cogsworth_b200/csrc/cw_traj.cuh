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
// completed a window and writes each of them cooperatively: 24 lanes x LDS.128 -> STG.128 = the six
// runs of one lane in ONE store instruction (12 full sectors, no partial-sector traffic).  The ragged
// first / last window of an orbit goes out as predicated 8-byte stores.
//
// (Round-1 note: the first version issued per-lane cp.async.bulk (TMA) stores.  UBLKCP is a
// uniform-datapath instruction, so ptxas serialised the 32 lanes with an R2UR/PLOP3/BRA loop -- 466
// warp instructions per orbit-step, 2.2 TB/s.  See profiles/r1_ncu_store_all_tma_per_lane_summary.md.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "cw_kernels.cuh"

namespace cw {

#ifndef CW_TRAJ_S
#define CW_TRAJ_S 8
#endif
constexpr int TRAJ_S = CW_TRAJ_S;                            // samples per window (4 or 8)
constexpr int TRAJ_CPR = TRAJ_S / 2;                         // 16-byte chunks per row
constexpr int TRAJ_LPS = 6 * TRAJ_CPR;                       // flush lanes per source lane (24 | 12)
constexpr int TRAJ_SPI = 32 / TRAJ_LPS;                      // source lanes per store instruction (1 | 2)
constexpr int TRAJ_ROWS = 6;                                 // x y z vx vy vz
constexpr int TRAJ_DATA_BYTES = TRAJ_ROWS * TRAJ_S * 8;      // 384
constexpr int TRAJ_LANE_BYTES = TRAJ_DATA_BYTES + 16;        // + header {col0, n}; also de-phases the banks

struct TrajHeader {
    long long col0; // absolute column of the first staged sample
    int n;          // staged samples
    int pad;
};

struct TrajWriter {
    unsigned char *warp_base; // the 32 lane regions of this warp
    double *buf;              // this lane's rows: buf[row * TRAJ_S + slot]
    double *my_row;           // flush role of this lane: destination row (lane >> 2), lanes 0..23
    int64_t col0;
    int n;
    bool want;                // my window is complete and waits for the next cooperative flush
    bool wide_ok;             // destination rows are 16-byte aligned -> 128-bit stores

    __device__ __forceinline__ void init(unsigned char *smem, const IntegArgs &A)
    {
        const int lane = threadIdx.x & 31;
        warp_base = smem + (size_t)(threadIdx.x - lane) * TRAJ_LANE_BYTES;
        buf = reinterpret_cast<double *>(warp_base + (size_t)lane * TRAJ_LANE_BYTES);
        const int r = min((lane % TRAJ_LPS) / TRAJ_CPR, TRAJ_ROWS - 1);
        my_row = (r < 3) ? A.traj_pos + (int64_t)r * A.traj_total : A.traj_vel + (int64_t)(r - 3) * A.traj_total;
        col0 = 0;
        n = 0;
        want = false;
        wide_ok = ((A.traj_total & 1) == 0) &&
                  (((reinterpret_cast<uintptr_t>(A.traj_pos) | reinterpret_cast<uintptr_t>(A.traj_vel)) & 15) == 0);
    }

    // stage one sample for absolute column `col`; `last` = final sample of the orbit.
    // At most one window may complete between two flush_warp() calls.
    __device__ __forceinline__ void put(int64_t col, double x, double y, double z, double vx, double vy, double vz,
                                        bool last)
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
        if (slot == TRAJ_S - 1 || last) {
            TrajHeader *h = reinterpret_cast<TrajHeader *>(reinterpret_cast<unsigned char *>(buf) + TRAJ_DATA_BYTES);
            h->col0 = col0;
            h->n = n;
            want = true;
        }
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

    // Warp-convergent: every lane of the warp must call this together.
    __device__ __forceinline__ void flush_warp()
    {
        unsigned m = __ballot_sync(0xffffffffu, want);
        if (m == 0) return;
        __syncwarp();                       // the staged samples / headers of the other lanes are visible
        const int lane = threadIdx.x & 31;
        const int sub = lane / TRAJ_LPS;                       // which source lane of a store instruction
        const int r = (lane % TRAJ_LPS) / TRAJ_CPR, c = lane % TRAJ_CPR;
        const bool role = lane < TRAJ_SPI * TRAJ_LPS;
        if (m == 0xffffffffu && wide_ok && __all_sync(0xffffffffu, n == TRAJ_S)) {
            // lockstep case (equal-length orbits): all 32 windows are complete -- fully unrolled, every
            // shared-memory offset an immediate: LDS.64 column, LDS.128 data, address, STG.128
            if (role) {
                const unsigned char *src = warp_base + sub * TRAJ_LANE_BYTES + r * (TRAJ_S * 8) + c * 16;
                const unsigned char *hdr = warp_base + sub * TRAJ_LANE_BYTES + TRAJ_DATA_BYTES;
                double *dst = my_row + 2 * c;
#pragma unroll
                for (int l = 0; l < 32; l += TRAJ_SPI) {
                    const long long c0 = *reinterpret_cast<const long long *>(hdr + l * TRAJ_LANE_BYTES);
                    const double2 v = *reinterpret_cast<const double2 *>(src + l * TRAJ_LANE_BYTES);
                    *reinterpret_cast<double2 *>(dst + c0) = v;
                }
            }
        } else {
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const unsigned char *src = warp_base + (size_t)l * TRAJ_LANE_BYTES;
                const TrajHeader h = *reinterpret_cast<const TrajHeader *>(src + TRAJ_DATA_BYTES);
                if (role && sub == 0) {
                    if (h.n == TRAJ_S && wide_ok) {
                        const double2 v = *reinterpret_cast<const double2 *>(src + r * (TRAJ_S * 8) + c * 16);
                        *reinterpret_cast<double2 *>(my_row + h.col0 + 2 * c) = v;
                    } else {
                        const int s0 = (int)(h.col0 & (TRAJ_S - 1));
                        const double *rowbuf = reinterpret_cast<const double *>(src + r * (TRAJ_S * 8));
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int e = 2 * c + q;
                            if (e >= s0 && e < s0 + h.n) my_row[h.col0 - s0 + e] = rowbuf[e];
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
