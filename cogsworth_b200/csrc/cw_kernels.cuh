// cw_kernels.cuh -- argument blocks and launch prototypes shared by the kernel files and the
// C-ABI host layer (cw_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "cw_device.cuh"

namespace cw {

// One device's shard of a cw_batch + the outputs of the grid pre-pass, all device pointers.
// Arrays are compact for the shard: component stride of w0 / w_final is `n`.
struct IntegArgs {
    int64_t n;        // orbits in this shard
    // inputs (cw_batch)
    const double *w0; // 6 x n
    const double *t1, *t2, *dt, *to_myr;
    const int32_t *flags;
    const int64_t *ev_off; // n + 1 (values carry the caller's base: subtract ev_base), may be null
    int64_t ev_base;
    const double *ev_time; // K
    const double *ev_dv;   // K x 3
    // grid pre-pass products
    int32_t *n_nodes;      // n
    int32_t *ev_step;      // K   node index of each kick
    double *ev_tk;         // K   accumulated grid time of that node (grid unit)
    double *t_last;        // n   accumulated grid time of the last accumulated node (may be null)
    int32_t *order;        // n   queue position -> orbit (longest first), may be null = identity
    unsigned long long *queue; // work counter of the persistent-lane scheduler
    // outputs (cw_result)
    double *w_final;       // 6 x n
    int32_t *status;       // n
    const int64_t *traj_off; // n + 1 (values carry the caller's base: subtract traj_base)
    int64_t traj_base;
    int64_t traj_total;    // component stride of traj_pos / traj_vel for this shard
    double *traj_pos, *traj_vel, *traj_t;
    // DOP853
    double atol, rtol;
    int64_t nmax;
    unsigned long long *stats; // 5: accepted, rejected, gradient evaluations, orbit-steps, orbits with status != OK
    int32_t *cost_key;         // n   DOP853 pilot: float bits of the estimated step attempts per orbit
    int32_t pilot_intervals;   // intervals the pilot integrates
};

struct LaunchCfg {
    int sm_count;
    cudaStream_t stream;
};

// grid pre-pass: node counts, kick node indices (bit-exact FP64 accumulation of the grid)
cudaError_t launch_grid_prepass(const IntegArgs &A, int32_t *iota_out, int32_t *sort_key_out, const LaunchCfg &cfg);
// fixed-step leapfrog (gala leapfrog_integrate_hamiltonian semantics)
cudaError_t launch_leapfrog(const PotParams &P, const IntegArgs &A, bool store_all, const LaunchCfg &cfg);
// orbits one full grid of the final-state leapfrog kernel of this potential kind holds at a time (SMs x resident CTAs x
// threads) on the CURRENT device; 0 on error.  Host-buffer calls cut their chunks in multiples of it.
int64_t leapfrog_lane_capacity(const PotParams &P, int sm_count);
// adaptive DOP853: restarted per grid interval (gala dop853_integrate_hamiltonian semantics), or
// dense = one call per segment with Hairer's continuous extension at the grid nodes
cudaError_t launch_dop853(const PotParams &P, const IntegArgs &A, bool store_all, bool dense, const LaunchCfg &cfg);
// DOP853 pilot: first A.pilot_intervals intervals of every orbit -> A.cost_key (queue ordering)
cudaError_t launch_dop853_pilot(const PotParams &P, const IntegArgs &A, const LaunchCfg &cfg);
// Orbit.t of stored trajectories (only when traj_t was requested)
cudaError_t launch_traj_times(const IntegArgs &A, const LaunchCfg &cfg);
// pointwise potential kernels (N3 rows); mode 0 gradient (3 x n), 1 value (n), 2 v_circ (n)
cudaError_t launch_pointwise(const PotParams &P, int mode, int64_t n, const double *q, const double *tt, double *out,
                             const LaunchCfg &cfg);
// diagnostics
cudaError_t launch_dfma_peak(double *sink, int iters, const LaunchCfg &cfg, int *blocks, int *threads);
cudaError_t launch_scatter_write(double2 *base, long long streams, long long stream_bytes, int run_bytes,
                                 int blocks, const LaunchCfg &cfg);
cudaError_t launch_math_selftest(int64_t n, double *out3, const LaunchCfg &cfg);

} // namespace cw
