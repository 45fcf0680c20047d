/*
 * cogsworth_b200.h -- C ABI of the B200-native galactic-orbit integrator.
 *
 * The reference (TomWagg/cogsworth) has no FFI seam on this path: the seam is Python,
 *     cogsworth.kicks.integrate_orbit_with_events(w0, t1, t2, dt, potential, events, store_all,
 *         integrator, integrator_kwargs, max_retries, timestep_multiplier)   src/cogsworth/kicks.py:52-56
 *     cogsworth.pop.Population.perform_galactic_evolution(progress_bar)      src/cogsworth/pop.py:1195
 *     cogsworth.pop._init_pool / _orbit_worker  (the multiprocessing boundary) src/cogsworth/pop.py:2187-2201
 * and underneath them gala's Cython operators
 *     leapfrog_integrate_hamiltonian(H, w0[norb,6], t[ntimes], store_all)
 *     dop853_integrate_hamiltonian(H, w0, t, atol, rtol, nmax, ...)
 * The entry points below are the BATCH form of that per-orbit call: one call integrates every
 * orbit that perform_galactic_evolution would fan out over its process pool.  The ctypes
 * binding a cogsworth maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / CUDA types (streams travel as void*).
 *   - all arrays are caller-owned; the library keeps no reference after a call returns.
 *   - phase-space arrays are structure-of-arrays, component-major:  w[c*n + i],
 *     c = 0..5 = (x, y, z [kpc], vx, vy, vz [kpc/Myr])   (gala's galactic unit system).
 *   - every entry point returns 0 or a negative CW_E_* code; per-orbit failures are reported in
 *     status[] with Hairer's DOP853 codes, which the Python layer maps to `None` orbits exactly
 *     as kicks.py:105-106,193-203 does.
 */
#ifndef COGSWORTH_B200_H
#define COGSWORTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CW_ABI_VERSION 2
#define CW_MAX_COMPONENTS 16

/* potential components = gala builtin potentials reachable from MilkyWayPotential v1 / v2
 * (gala/potential/potential/builtin/builtin_potentials.c); params are 3 doubles per component */
enum {
    CW_COMP_MIYAMOTO_NAGAI = 1, /* (m [Msun], a [kpc], b [kpc])  -- MN3ExponentialDisk = three of these */
    CW_COMP_HERNQUIST = 2,      /* (m, c, unused) */
    CW_COMP_NFW_SPHERICAL = 3,  /* (m, r_s, unused) */
    /* further spherical / axisymmetric gala builtins for user-supplied CompositePotentials (generic kernel) */
    CW_COMP_KEPLER = 4,         /* (m, unused, unused) */
    CW_COMP_PLUMMER = 5,        /* (m, b, unused) */
    CW_COMP_ISOCHRONE = 6,      /* (m, b, unused) */
    CW_COMP_JAFFE = 7,          /* (m, c, unused) */
    CW_COMP_LOGARITHMIC = 8,    /* (v_c [kpc/Myr], r_h, q3) with q1 = q2 = 1 */
    CW_COMP_KUZMIN = 9,         /* (m, a, unused) */
    CW_COMP_BURKERT = 10,       /* (rho [Msun/kpc^3], r0, unused) */
    CW_COMP_NFW_TRIAXIAL = 11,  /* (m, r_s, a, b, c): gp.NFWPotential with flattening; cw_set_potential_ex only */
    CW_COMP_LOG_TRIAXIAL = 12,  /* (v_c, r_h, q1, q2, q3): LogarithmicPotential with axis ratios, phi = 0; _ex only */
    /* components whose gradient is not of the form (Fx x, Fy y, Fz z): cw_set_potential_general only */
    CW_COMP_LONG_MURALI_BAR = 13, /* (m, a, b, c): gp.LongMuraliBarPotential; its angle alpha travels as a rotation */
    CW_COMP_LEESUTO_NFW = 14,     /* (v_c [kpc/Myr], r_s, a, b, c): gp.LeeSutoTriaxialNFWPotential */
    CW_COMP_POWERLAW_CUTOFF = 15  /* (m, alpha, r_c): gp.PowerLawCutoffPotential (BovyMWPotential2014's bulge) */
};
#define CW_MAX_KNOTS 64 /* time knots of all components of one potential together */
enum { CW_INTERP_LINEAR = 0, CW_INTERP_CUBIC = 1 /* natural cubic spline (GSL cspline) */ };

/* integrators = the two gala Integrator classes cogsworth passes (pop.py:130, kicks.py:54) */
enum {
    CW_LEAPFROG = 0, /* gi.LeapfrogIntegrator : fixed-step KDK, 1 gradient per grid interval */
    CW_DOP853 = 1,   /* gi.DOPRI853Integrator : Hairer dop853 restarted on every grid interval */
    CW_DOP853_DENSE = 2 /* gi.DOPRI853Integrator, dense-output form: ONE dop853 call per segment between
                           kicks (h0 = t[1]-t[0], hmax = the segment), grid nodes inside a step evaluated with
                           Hairer's continuous extension contd8; the segment's end state is the integrator's own */
};

/* per-orbit grid flags */
enum {
    CW_GRID_APPEND_T2 = 1 /* kicks.py:121-122: append t2 as the final node (events branch; backward grids) */
};

/* buffer placement */
enum {
    CW_MEM_HOST = 0,  /* every pointer in cw_batch / cw_result is host memory (page-locked or pageable, see cw_host_alloc) */
    CW_MEM_DEVICE = 1 /* every pointer is device memory on the context's FIRST device */
};

/* call-level error codes */
enum {
    CW_OK = 0,
    CW_E_INVALID = -1,      /* bad argument */
    CW_E_CUDA = -2,         /* CUDA runtime error (see cw_last_error) */
    CW_E_NO_POTENTIAL = -3, /* cw_set_potential not called */
    CW_E_NO_DEVICE = -4,    /* no usable sm_100 device: there is no CPU fallback */
    CW_E_UNSUPPORTED = -5
};

/* per-orbit status (Hairer dop853 codes; -5 and -6 are ours) */
enum {
    CW_ORBIT_OK = 0,
    CW_ORBIT_INCONSISTENT = -1,    /* empty grid, unsorted events, or an event after t2 (the reference
                                      raises ValueError when stitching, kicks.py:187-189) */
    CW_ORBIT_NMAX = -2,            /* "Larger nmax is needed" */
    CW_ORBIT_STEP_TOO_SMALL = -3,  /* "Step size becomes too small" */
    CW_ORBIT_STIFF = -4,           /* "The problem is probably stiff" */
    CW_ORBIT_NONFINITE = -5,       /* error norm became NaN/inf (Hairer would spin to nmax) */
    CW_ORBIT_CAPACITY = -6         /* trajectory slot smaller than the orbit's node count */
};

typedef struct cw_ctx cw_ctx;

/* Options of one integrate call = the static arguments of pop._init_pool (pop.py:2187-2194). */
typedef struct cw_opts {
    int32_t integrator; /* CW_LEAPFROG | CW_DOP853 | CW_DOP853_DENSE */
    int32_t mem_space;  /* CW_MEM_HOST | CW_MEM_DEVICE */
    double atol;        /* integrator_kwargs["atol"], gala default 1e-10 (DOP853 only) */
    double rtol;        /* integrator_kwargs["rtol"], gala default 1e-10 */
    int64_t nmax;       /* integrator_kwargs["nmax"], <= 0 -> 100000 */
    void *stream;       /* CW_MEM_DEVICE only: cudaStream_t to launch on (NULL = the context's own) */
    int32_t no_sync;    /* CW_MEM_DEVICE only: return without synchronising the stream.  A context has ONE set of
                           scratch buffers: the next call on the same context (any stream) first waits, on the device,
                           for the previous no_sync call to finish; results are valid once `stream` is synchronised */
    int32_t reserved;
} cw_opts;

/*
 * The orbits of one call = what pop._iter_orbit_args yields (pop.py:1156-1193), flattened.
 *
 * Time grid of orbit i (kicks.py:97,118-122 and gala.integrate.parse_time_specification):
 *     T_0 = t1[i];  T_{j+1} = T_j + dt[i]  while T_j < t2[i]  (FP64 accumulation, <= 1e6 nodes);
 *     if (grid_flags[i] & CW_GRID_APPEND_T2) and T_last < t2[i]:  append t2[i].
 * Backward grids (t2 < t1 with dt < 0; cogsworth hydro/rewind.py:16-19): T_{j+1} = T_j + dt while T_j > t2, then t2
 * itself when CW_GRID_APPEND_T2 is set (gala's backward branch); such orbits must not carry events.
 * t1/t2/dt/ev_time are in the orbit's own GRID UNIT and to_myr[i] converts it to Myr by
 * multiplication (cogsworth builds the grid in seconds on the events branch -- to_myr =
 * 1/3.15576e13 -- and gala builds it in Myr on the no-events branch -- to_myr = 1).  The caller
 * must already have applied dt = min(dt, t2 - t1) (kicks.py:97).
 *
 * Events of orbit i are rows ev_off[i] .. ev_off[i+1]-1, chronological.  Event e is applied at
 * the last node k >= cursor with T_k < ev_time[e] (kicks.py:134); the velocity change ev_dv[e]
 * is get_kick_differential(...) (kicks.py:12-49) already rotated and converted to kpc/Myr.
 */
/* One class of time grid: the four per-orbit grid columns of the plain layout.  A population has two of them --
 * the events branch (seconds, t2 appended) and gala's own grid (Myr) -- plus one for every orbit whose span is
 * shorter than dt (kicks.py:97), so the columns compress to one byte per orbit. */
typedef struct cw_grid_class {
    double t2, dt, to_myr;
    int32_t grid_flags;
    int32_t reserved;
} cw_grid_class;

typedef struct cw_batch {
    int64_t n_orbits;
    const double *w0;          /* 6 x n_w0 (n_w0 = n_orbits in the plain layout) */
    const double *t1;          /* n_w0 */
    const double *t2;          /* n_orbits   (NULL when grid_class is given) */
    const double *dt;          /* n_orbits   (NULL when grid_class is given) */
    const double *to_myr;      /* n_orbits   (NULL when grid_class is given) */
    const int32_t *grid_flags; /* n_orbits   (NULL when grid_class is given) */
    const int64_t *ev_off;     /* n_orbits + 1, or NULL when no orbit has events (or when ev_off32 is given) */
    const double *ev_time;     /* K = ev_off[n_orbits] */
    const double *ev_dv;       /* K x 3 (row-major: dvx, dvy, dvz per event) */
    /* ---- ABI 2: compact forms, CW_MEM_HOST only; all zero = the plain layout above.  They exist to cut the bytes
     * that cross the host link (134 -> 78 B per orbit on a kicked population); a small kernel expands them on the
     * device, so the integrators see the plain layout either way and return the same bits. */
    int64_t n_w0;              /* 0 -> n_orbits.  Otherwise orbits [0, n_w0) start from column i of w0 / t1 and orbit
                                  n_w0 + k from column w0_index[k]: disrupted secondaries re-use their primary's
                                  initial conditions and birth time (pop.py:1186-1193) */
    const int32_t *w0_index;   /* n_orbits - n_w0 entries, non-decreasing, each in [0, n_w0) */
    const uint8_t *grid_class; /* n_orbits: index into classes[] */
    const cw_grid_class *classes;
    int32_t n_classes;         /* <= 256 */
    int32_t reserved;
    const int32_t *ev_off32;   /* n_orbits + 1: ev_off in 32 bits (K < 2^31) */
} cw_batch;

/* Outputs.  Any pointer except w_final and status may be NULL. */
typedef struct cw_result {
    double *w_final;   /* 6 x n_orbits : state at the last grid node = orbit[-1] (kicks.py:108-110,206-207) */
    int32_t *status;   /* n_orbits     : CW_ORBIT_* */
    int32_t *n_nodes;  /* n_orbits     : number of grid nodes = len(orbit.t) */
    int32_t *ev_step;  /* K            : node index at which each kick was applied (bit-exact check) */
    /* full trajectories (store_all=True), in the layout Population.save writes (pop.py:1853-1876):
       sample j of orbit i lives at column traj_off[i] + j */
    const int64_t *traj_off; /* n_orbits + 1, or NULL for final-state-only runs */
    double *traj_pos;        /* 3 x traj_off[n_orbits]  kpc */
    double *traj_vel;        /* 3 x traj_off[n_orbits]  kpc/Myr */
    double *traj_t;          /* traj_off[n_orbits]      Myr (may be NULL) */
    int64_t *stats;          /* 4 : DOP853 accepted steps, rejected steps, gradient evaluations, orbit-steps */
    int64_t *n_failed;       /* 1 : number of orbits whose status is not CW_ORBIT_OK (ABI 2; may be NULL) -- lets the
                                caller skip every per-orbit scan of a clean batch */
} cw_result;

/* Context: owns streams and scratch on each device.  n_devices <= 0 -> every visible device.
 * Replaces Population._ensure_pool / _close_pool (pop.py:917-939). */
int cw_create(cw_ctx **ctx, int n_devices, const int *device_ids);
void cw_destroy(cw_ctx *ctx);
int cw_device_count(const cw_ctx *ctx);

/* Replaces the `potential` static argument of _init_pool (pop.py:2188): an ordered list of
 * analytic components in galactic units; G in kpc^3 / Msun / Myr^2. */
int cw_set_potential(cw_ctx *ctx, double G, int n_components, const int32_t *types, const double *params);
/* n_components = 0 is gala's NullPotential (free motion). */
/* The same with `stride` (3..6) doubles per component, for components with more than three parameters. */
int cw_set_potential_ex(cw_ctx *ctx, double G, int n_components, const int32_t *types, const double *params, int stride);
/* The general form (its own kernels: full gradient vector, time argument).  Adds, per component,
 *   rotations : NULL, or n_components x 9 doubles, row-major R with q' = R q evaluated in the component's
 *               frame and grad = R^T grad' (gala's `R=` argument of every potential class; the `phi` of
 *               LogarithmicPotential / LM10Potential and the `alpha` of LongMuraliBarPotential).  A row whose
 *               first entry is NaN means "no rotation".
 *   knot_off  : NULL, or n_components + 1 offsets (CSR) into knot_t / knot_amp: gp.TimeInterpolatedPotential
 *               (docs/tutorials/potentials/evolving_potentials.ipynb cells 5 and 13) with a time-dependent first
 *               parameter (the mass; the other parameters must be constant).  knot_t [Myr] strictly increasing,
 *               knot_amp = the first parameter at the knots; `interp` = CW_INTERP_LINEAR / CW_INTERP_CUBIC.
 *               Outside the knots the end values are held.  A component with 0 knots is static.
 * Potentials that need none of this and contain no non-diagonal component are routed to the same kernels as
 * cw_set_potential_ex.  When every component carries the same knot times and every amplitude curve is the same
 * multiple of its own last value (one mass-fraction array for all components, as in the tutorial), the potential is
 * separable, Phi(q, t) = f(t) Phi_0(q), and runs in the static kernels with one scale factor per gradient. */
int cw_set_potential_general(cw_ctx *ctx, double G, int n_components, const int32_t *types, const double *params,
                             int stride, const double *rotations, const int32_t *knot_off, const double *knot_t,
                             const double *knot_amp, int interp);
/* Time [Myr] at which cw_gradient / cw_potential_value / cw_circular_velocity evaluate a time-dependent potential
 * (default 0; the integrators use each orbit's own clock). */
int cw_set_potential_time(cw_ctx *ctx, double t_myr);

/* The grid pre-pass alone: node count per orbit and the node index of every event.  Lets the
 * host size trajectory buffers (traj_off) before cw_integrate, and is the bit-exact
 * kick-timing check on its own. */
int cw_count_nodes(cw_ctx *ctx, const cw_opts *opts, const cw_batch *batch, int32_t *n_nodes, int32_t *ev_step);
/* batch->w0 may be NULL here (the grid does not depend on it), and with ev_step == NULL the events are not read
 * either: sizing the trajectory buffers of a store_all run moves 28 B per orbit (1 B with grid_class). */

/* The hot path: replaces pool.starmap(_orbit_worker, ...) (pop.py:1245-1266).  Orbits are sharded
 * by index over the context's devices; there is no collective. */
int cw_integrate(cw_ctx *ctx, const cw_opts *opts, const cw_batch *batch, const cw_result *result);

/* "Next" rows either side of the path: potential gradient / value / circular
 * velocity for n points, q is 3 x n SoA.  Replaces potential.gradient / .energy /
 * .circular_velocity as used at pop.py:886-895,1001-1013. */
int cw_gradient(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *grad /* 3 x n */);
int cw_potential_value(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *phi /* n */);
int cw_circular_velocity(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *vcirc /* n, kpc/Myr */);
/* The same three with a time per point [Myr] (t_myr = NULL: the time of cw_set_potential_time): what a
 * time-dependent potential needs for `galactic_potential.circular_velocity(q, t = max_ev_time - tau)` at every
 * system's own birth time (pop.py:1001-1004).  what: CW_GRADIENT (out 3 x n), CW_VALUE (n), CW_VCIRC (n). */
enum { CW_GRADIENT = 0, CW_VALUE = 1, CW_VCIRC = 2 };
int cw_pointwise_t(cw_ctx *ctx, int what, int mem_space, int64_t n, const double *q, const double *t_myr, double *out);

/* Page-locked host memory from the library's cache (cudaHostAlloc, portable; freed blocks are kept for re-use up to
 * CW_PINNED_CACHE_MB, default 4096).  Host buffers handed to cw_integrate may be ANY host memory: page-locked buffers
 * (these, torch pinned tensors, cudaHostRegister'ed arrays) are copied from / to directly; pageable ones (plain
 * numpy) are staged through the library's own page-locked ring by a pool of host threads (CW_HOST_THREADS), chunk by
 * chunk, overlapped with the kernels.  Replaces the pickling of arguments / results across the process pool
 * (pop.py:1245-1266). */
void *cw_host_alloc(size_t bytes);
void cw_host_free(void *p);
/* A page-locked block that is ALREADY in the library's cache, or NULL -- never pins new memory.  cudaHostAlloc costs
 * 0.45 ms per MB (measured: 216 ms for the 480 MB of final states of 10^7 orbits) against 3 ms per call that DMA from
 * page-locked arrays saves over the staging ring, so a caller that runs the stage ONCE (Population.
 * perform_galactic_evolution) asks here and falls back to ordinary memory; a caller in a loop uses cw_host_alloc. */
void *cw_host_alloc_cached(size_t bytes);
/* Give every cached (free) page-locked block back to the driver; blocks in use are not touched. */
void cw_host_trim(void);
/* 1 if p is page-locked (or device) memory known to the CUDA runtime, 0 if pageable */
int cw_host_is_pinned(const void *p);
/* get_kick_differential (kicks.py:12-49) for K events at once, on the library's host threads: dv (K x 3, km/s in the
 * BSE frame), phase, inc [rad] -> out (K x 3) = R(phase, inc) dv * scale (scale = 1/977.7922216807891 gives the
 * kpc/Myr that cw_batch.ev_dv wants).  Host-side preparation of the batch, not part of the timed hot path. */
void cw_host_rotate_kicks(int64_t K, const double *dv, const double *phase, const double *inc, double scale, double *out);

/* Diagnostics */
const char *cw_strerror(int code);
const char *cw_last_error(const cw_ctx *ctx);
int cw_abi_version(void);
/* kernels launched by the last cw_integrate / cw_count_nodes call, over all devices */
int64_t cw_last_launch_count(const cw_ctx *ctx);
/* DOPRI853: the largest number of step attempts any ONE orbit of the last synchronous call needed.  An orbit is
 * sequential, so attempts x (latency of one attempt) of this orbit is a lower bound of the launch's duration. */
int64_t cw_last_max_attempts(const cw_ctx *ctx);
/* device time (ms, CUDA events on the launching stream) of the integrator kernel / of all kernels
 * of the last call on device slot `dev` */
double cw_last_kernel_ms(const cw_ctx *ctx, int dev);
double cw_last_total_kernel_ms(const cw_ctx *ctx, int dev);
/* DFMA micro-benchmark on device slot `dev`: returns sustained FP64 TFLOP/s (2 flop per FMA) over
 * about `ms_target` milliseconds -- the roofline denominator for this path. */
double cw_measure_fp64_peak(cw_ctx *ctx, int dev, double ms_target);
/* Scatter-write micro-benchmark = the HBM access pattern of store_all without the arithmetic: `streams`
 * far-apart sequential streams of `stream_bytes`, appended `run_bytes` (16..512) at a time in lock-step by
 * `resident_threads_per_sm` threads per SM.  Returns GB/s written (the pattern's ceiling). */
double cw_measure_scatter_write(cw_ctx *ctx, int dev, int64_t streams, int64_t stream_bytes, int run_bytes,
                                int resident_threads_per_sm);
/* max relative error of the kernels' own rsqrt / reciprocal / log (table-free) / log (table)
 * against the CUDA math library over n samples (out[0..3]), then the dependent-issue latency in cycles of
 * DFMA, rsqrt_fast + DADD and DMUL measured by one warp alone (out[4..6]); out[7] */
int cw_selftest_math(cw_ctx *ctx, int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif /* COGSWORTH_B200_H */
