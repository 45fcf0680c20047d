// cw_api.cu -- the C ABI (include/cogsworth_b200.h): context, device memory, sharding of a batch
// over the context's devices, kernel sequencing.  This is the B200 replacement for the
// multiprocessing boundary of the reference (cogsworth pop.py:917-939 _ensure_pool,
// :1240-1266 pool.starmap(_orbit_worker), :2187-2201 _init_pool/_orbit_worker).
//
// There is NO CPU fallback anywhere in this file: without a CUDA device cw_create fails with
// CW_E_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <climits>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>      // header-only: ranges show up in Nsight Systems / Compute timelines, cost nothing otherwise

#include "cw_host.h"
#include "cw_kernels.cuh"

using namespace cw;

namespace {

// restores the caller's current device on every exit path (the CW_CK macro returns early on CUDA errors)
// NVTX range for the host-side timeline (tracing): cw_integrate / per-slot thread / per-chunk enqueue / drain
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct DeviceGuard {
    int prev = 0;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaError_t reserve(size_t bytes)
    {
        used = 0;
        if (bytes <= cap) return cudaSuccess;
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + (1u << 20);
        cudaError_t e = cudaMalloc((void **)&base, want);
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    template <class T> T *take(size_t count)
    {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (used + bytes > cap) return nullptr;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
    static size_t padded(size_t count, size_t elt) { return (count * elt + 255) & ~size_t(255); }
    void release()
    {
        if (base) cudaFree(base);
        base = nullptr;
        cap = used = 0;
    }
};

struct Slot {
    static constexpr int N_LANES = 3;
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;          // = lanes[0]
    cudaStream_t lanes[N_LANES] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;   // diagnostics entry points
    std::vector<cudaEvent_t> events;        // per-chunk events (grow-only pool)
    Arena arena;                            // diagnostics / pointwise entry points
    // Integration scratch, ONE ARENA PER STREAM LANE: the chunks of a lane run one after another on the lane's
    // stream, so each of them re-uses the lane's buffers (stream order is the only synchronisation needed).
    // Device memory of a host-buffer call is therefore bounded by N_LANES chunks, not by the size of the batch:
    // trajectories larger than HBM stream through.
    Arena lane_arena[N_LANES];
    Arena meta;                             // queue counter + statistics of every chunk of a call (one padded block
                                            // each): NOT re-used within a call, so they can be read back at its end
    // page-locked staging buffers of the lanes (from the library's cache) for pageable caller arrays
    void *stage_in[N_LANES] = {nullptr, nullptr, nullptr};
    void *stage_out[N_LANES] = {nullptr, nullptr, nullptr};
    size_t stage_in_cap[N_LANES] = {0, 0, 0}, stage_out_cap[N_LANES] = {0, 0, 0};
    cudaEvent_t h2d_done[N_LANES] = {nullptr, nullptr, nullptr};   // the DMA out of stage_in[l] has completed
    bool h2d_pending[N_LANES] = {false, false, false};
    // device-mode no_sync calls: the next call is ordered behind this event (they share the scratch above)
    cudaEvent_t pending = nullptr;
    bool has_pending = false;
    double last_kernel_ms = 0.0, last_total_ms = 0.0;
    int64_t launches = 0;                   // kernels launched by the current call on this slot
    std::string error;                      // failure text of the slot's worker thread
    cudaError_t ensure_events(size_t n)
    {
        while (events.size() < n) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
            events.push_back(e);
        }
        return cudaSuccess;
    }
};

} // namespace

// One long-lived host thread per device slot beyond the first (the caller's thread drives the first busy slot): a
// host-mode call hands each its slot's work and waits.  (Spawning the threads per call cost 0.3 ms of a 10 ms call on 8
// GPUs.)
struct SlotWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, stop = false;

    void start()
    {
        th = std::thread([this] {
            for (;;) {
                std::function<void()> j;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || has_job; });
                    if (stop) return;
                    j = std::move(job);
                    has_job = false;
                }
                j();
                {
                    std::lock_guard<std::mutex> lk(m);
                    done = true;
                }
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> f)
    {
        if (!th.joinable()) start();
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(f);
            has_job = true;
            done = false;
        }
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
    }
    ~SlotWorker()
    {
        if (th.joinable()) {
            {
                std::lock_guard<std::mutex> lk(m);
                stop = true;
            }
            cv.notify_all();
            th.join();
        }
    }
};

struct cw_ctx {
    std::vector<Slot> slots;
    std::vector<std::unique_ptr<SlotWorker>> workers;   // [d] serves slot d (entry 0 stays empty until slot 0 is not the first busy one)
    PotParams pot;
    bool has_pot = false;
    std::string last_error;
    int64_t last_launches = 0;
    int64_t last_max_attempts = 0;          // DOP853: most step attempts of any one orbit in the last call
};

#define CW_CK(ctx, call)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (ctx)->last_error = _b;                                                               \
            return CW_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

static int fail(cw_ctx *ctx, int code, const char *msg)
{
    if (ctx) ctx->last_error = msg;
    return code;
}

extern "C" int cw_abi_version(void) { return CW_ABI_VERSION; }

extern "C" const char *cw_strerror(int code)
{
    switch (code) {
    case CW_OK: return "ok";
    case CW_E_INVALID: return "invalid argument";
    case CW_E_CUDA: return "CUDA runtime error";
    case CW_E_NO_POTENTIAL: return "no potential set (call cw_set_potential)";
    case CW_E_NO_DEVICE: return "no CUDA device available (this library has no CPU fallback)";
    case CW_E_UNSUPPORTED: return "unsupported";
    default: return "unknown error";
    }
}

extern "C" const char *cw_last_error(const cw_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int cw_create(cw_ctx **out, int n_devices, const int *device_ids)
{
    if (!out) return CW_E_INVALID;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
        cudaGetLastError();
        return CW_E_NO_DEVICE;
    }
    cw_ctx *ctx = new (std::nothrow) cw_ctx();
    if (!ctx) return CW_E_INVALID;
    memset(&ctx->pot, 0, sizeof ctx->pot);
    std::vector<int> ids;
    if (n_devices <= 0) {
        for (int d = 0; d < visible; d++) ids.push_back(d);
    } else {
        for (int k = 0; k < n_devices; k++) {
            int d = device_ids ? device_ids[k] : k;
            if (d < 0 || d >= visible) {
                delete ctx;
                return CW_E_INVALID;
            }
            ids.push_back(d);
        }
    }
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d : ids) {
        Slot s;
        s.dev = d;
        cudaError_t e = cudaSetDevice(d);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, d);
        for (int l = 0; l < Slot::N_LANES && e == cudaSuccess; l++)
            e = cudaStreamCreateWithFlags(&s.lanes[l], cudaStreamNonBlocking);
        s.stream = s.lanes[0];
        if (e == cudaSuccess) e = cudaEventCreate(&s.ev_t0);
        if (e == cudaSuccess) e = cudaEventCreate(&s.ev_k0);
        if (e == cudaSuccess) e = cudaEventCreate(&s.ev_k1);
        for (int l = 0; l < Slot::N_LANES && e == cudaSuccess; l++)
            e = cudaEventCreateWithFlags(&s.h2d_done[l], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.pending, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaSetDevice(prev);
            delete ctx;
            return CW_E_CUDA;
        }
        ctx->slots.push_back(s);
    }
    cudaSetDevice(prev);
    *out = ctx;
    return CW_OK;
}

extern "C" void cw_destroy(cw_ctx *ctx)
{
    if (!ctx) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (Slot &s : ctx->slots) {
        cudaSetDevice(s.dev);
        for (int l = 0; l < Slot::N_LANES; l++)
            if (s.lanes[l]) cudaStreamSynchronize(s.lanes[l]);
        s.arena.release();
        for (int l = 0; l < Slot::N_LANES; l++) s.lane_arena[l].release();
        s.meta.release();
        for (cudaEvent_t e : s.events) cudaEventDestroy(e);
        if (s.ev_t0) cudaEventDestroy(s.ev_t0);
        if (s.ev_k0) cudaEventDestroy(s.ev_k0);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.has_pending) cudaEventSynchronize(s.pending);
        if (s.pending) cudaEventDestroy(s.pending);
        for (int l = 0; l < Slot::N_LANES; l++) {
            if (s.h2d_done[l]) cudaEventDestroy(s.h2d_done[l]);
            pinned_free(s.stage_in[l]);
            pinned_free(s.stage_out[l]);
        }
        for (int l = 0; l < Slot::N_LANES; l++)
            if (s.lanes[l]) cudaStreamDestroy(s.lanes[l]);
    }
    cudaSetDevice(prev);
    delete ctx;
}

extern "C" int cw_device_count(const cw_ctx *ctx) { return ctx ? (int)ctx->slots.size() : 0; }

extern "C" int cw_set_potential_ex(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params_in, int stride);

extern "C" int cw_set_potential(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params)
{
    return cw_set_potential_ex(ctx, G, n_comp, types, params, 3);
}

static int set_potential_impl(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params_in, int stride,
                              const double *rotations, const int32_t *knot_off, const double *knot_t,
                              const double *knot_amp, int interp);

extern "C" int cw_set_potential_ex(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params_in, int stride)
{
    return set_potential_impl(ctx, G, n_comp, types, params_in, stride, nullptr, nullptr, nullptr, nullptr, 0);
}

extern "C" int cw_set_potential_general(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params_in,
                                        int stride, const double *rotations, const int32_t *knot_off,
                                        const double *knot_t, const double *knot_amp, int interp)
{
    return set_potential_impl(ctx, G, n_comp, types, params_in, stride, rotations, knot_off, knot_t, knot_amp, interp);
}

extern "C" int cw_set_potential_time(cw_ctx *ctx, double t_myr)
{
    if (!ctx) return CW_E_INVALID;
    ctx->pot.t_eval = t_myr;
    return CW_OK;
}

// piecewise-cubic coefficients of the interpolant through (t_j, y_j), j = 0..n-1, on interval j in powers of (t - t_j):
// CW_INTERP_LINEAR, or the natural cubic spline (second derivative zero at both ends; GSL's gsl_interp_cspline)
static void spline_coefficients(int n, const double *t, const double *y, int interp, double (*c)[4])
{
    std::vector<double> M(n, 0.0);
    if (interp == CW_INTERP_CUBIC && n > 2) {
        // tridiagonal system for the interior second derivatives (Thomas algorithm)
        std::vector<double> cp(n, 0.0), dp(n, 0.0);
        for (int j = 1; j < n - 1; j++) {
            const double h0 = t[j] - t[j - 1], h1 = t[j + 1] - t[j];
            const double diag = 2.0 * (h0 + h1);
            const double rhs = 6.0 * ((y[j + 1] - y[j]) / h1 - (y[j] - y[j - 1]) / h0);
            const double den = diag - h0 * cp[j - 1];
            cp[j] = h1 / den;
            dp[j] = (rhs - h0 * dp[j - 1]) / den;
        }
        for (int j = n - 2; j >= 1; j--) M[j] = dp[j] - cp[j] * M[j + 1];
    }
    for (int j = 0; j < n; j++) {
        c[j][0] = y[j]; c[j][1] = c[j][2] = c[j][3] = 0.0;
        if (j + 1 < n) {
            const double h = t[j + 1] - t[j];
            c[j][1] = (y[j + 1] - y[j]) / h - h * (2.0 * M[j] + M[j + 1]) / 6.0;
            c[j][2] = 0.5 * M[j];
            c[j][3] = (M[j + 1] - M[j]) / (6.0 * h);
        }
    }
}

static int set_potential_impl(cw_ctx *ctx, double G, int n_comp, const int32_t *types, const double *params_in, int stride,
                              const double *rotations, const int32_t *knot_off, const double *knot_t,
                              const double *knot_amp, int interp)
{
    if (stride < 3 || stride > 6) return fail(ctx, CW_E_INVALID, "cw_set_potential_ex: stride must be 3..6");
    // the three leading parameters in the layout the rest of this function reads
    double params[3 * CW_MAX_COMPONENTS] = {0};
    if (n_comp > 0 && n_comp <= CW_MAX_COMPONENTS && params_in)
        for (int i = 0; i < n_comp; i++)
            for (int k = 0; k < 3; k++) params[3 * i + k] = params_in[(size_t)stride * i + k];
    if (!ctx || n_comp < 0 || n_comp > CW_MAX_COMPONENTS || (n_comp > 0 && (!types || !params_in)))
        return fail(ctx, CW_E_INVALID, "cw_set_potential: bad arguments");
    PotParams P;
    memset(&P, 0, sizeof P);
    P.n_comp = n_comp;
    P.G = G;
    bool general = false, separable = false;
    // Separable time dependence: every component carries knots at the same times and every amplitude curve is the
    // same multiple of its own last value (the tutorial's `mass * mass_fractions`), nothing is rotated or
    // non-diagonal.  Then Phi(q, t) = f(t) Phi_0(q) and the static kernels run with one scale factor.
    if (knot_off && n_comp > 0 && knot_t && knot_amp && knot_off[0] == 0 && knot_off[1] >= 2 && knot_off[1] <= CW_MAX_KNOTS &&
        (interp == CW_INTERP_LINEAR || interp == CW_INTERP_CUBIC)) {
        const int nk = knot_off[1];
        separable = true;
        for (int i = 0; i < n_comp && separable; i++) {
            const int t = types[i];
            if (knot_off[i + 1] - knot_off[i] != nk || t >= CW_COMP_LONG_MURALI_BAR || t == CW_COMP_LOGARITHMIC ||
                t == CW_COMP_LOG_TRIAXIAL || t < CW_COMP_MIYAMOTO_NAGAI) { separable = false; break; }
            const double *ti = knot_t + knot_off[i], *ai = knot_amp + knot_off[i];
            const double ref_i = ai[nk - 1], ref_0 = knot_amp[nk - 1];
            if (!(std::isfinite(ref_i) && ref_i != 0.0 && ref_0 != 0.0)) { separable = false; break; }
            for (int j = 0; j < nk; j++) {
                const double r_i = ai[j] / ref_i, r_0 = knot_amp[j] / ref_0;
                if (!(ti[j] == knot_t[j]) || !std::isfinite(ai[j]) || fabs(r_i - r_0) > 4e-16 * fmax(fabs(r_0), 1.0) ||
                    (j > 0 && !(ti[j] > ti[j - 1]))) { separable = false; break; }
            }
        }
        if (rotations)
            for (int i = 0; i < n_comp && separable; i++) {
                const double *Rm = rotations + 9 * (size_t)i;
                bool identity = true;
                for (int k = 0; k < 9; k++) identity = identity && (Rm[k] == ((k % 4 == 0) ? 1.0 : 0.0));
                if (!std::isnan(Rm[0]) && !identity) separable = false;
            }
        if (separable) {
            std::vector<double> ratio(nk);
            for (int j = 0; j < nk; j++) { ratio[j] = knot_amp[j] / knot_amp[nk - 1]; P.k_t[j] = knot_t[j]; }
            spline_coefficients(nk, knot_t, ratio.data(), interp, P.k_c);
            P.k_off[0] = 0;
            for (int i = 1; i <= n_comp; i++) P.k_off[i] = nk;      // f(t) lives in the slot of "component 0"
            for (int i = 0; i < n_comp; i++) params[3 * i + 0] = knot_amp[knot_off[i] + nk - 1];
        }
    }
    if (knot_off && !separable) {
        if (interp != CW_INTERP_LINEAR && interp != CW_INTERP_CUBIC)
            return fail(ctx, CW_E_INVALID, "cw_set_potential_general: interp must be CW_INTERP_LINEAR or CW_INTERP_CUBIC");
        if (knot_off[0] != 0 || knot_off[n_comp] > CW_MAX_KNOTS)
            return fail(ctx, CW_E_INVALID, "cw_set_potential_general: more than CW_MAX_KNOTS knots, or knot_off[0] != 0");
        for (int i = 0; i < n_comp; i++) {
            const int a = knot_off[i], b = knot_off[i + 1];
            if (b < a) return fail(ctx, CW_E_INVALID, "cw_set_potential_general: knot_off must not decrease");
            P.k_off[i] = a;
            if (b == a) continue;
            if (!knot_t || !knot_amp) return fail(ctx, CW_E_INVALID, "cw_set_potential_general: knots without times / amplitudes");
            const int t = types[i];
            if (t == CW_COMP_LOGARITHMIC || t == CW_COMP_LOG_TRIAXIAL || t == CW_COMP_LEESUTO_NFW)
                return fail(ctx, CW_E_UNSUPPORTED, "cw_set_potential_general: a time-dependent v_c is not supported (masses and densities are)");
            for (int j = a; j < b; j++) {
                if (!std::isfinite(knot_t[j]) || !std::isfinite(knot_amp[j]) || (j > a && !(knot_t[j] > knot_t[j - 1])))
                    return fail(ctx, CW_E_INVALID, "cw_set_potential_general: knot times must be finite and strictly increasing");
                P.k_t[j] = knot_t[j];
            }
            spline_coefficients(b - a, knot_t + a, knot_amp + a, interp, P.k_c + a);
            params[3 * i + 0] = 1.0;
            general = true;
        }
        P.k_off[n_comp] = knot_off[n_comp];
    }
    // equally spaced knots (np.linspace) are indexed directly
    for (int i = 0; i < n_comp; i++) {
        const int a = P.k_off[i], b = P.k_off[i + 1];
        if (b - a < 3) continue;
        const double dt = (P.k_t[b - 1] - P.k_t[a]) / (double)(b - a - 1);
        bool uniform = true;
        for (int j = a; j < b; j++) uniform = uniform && fabs(P.k_t[j] - (P.k_t[a] + dt * (double)(j - a))) <= 1e-12 * fabs(dt);
        if (uniform) { P.k_t0[i] = P.k_t[a]; P.k_inv_dt[i] = 1.0 / dt; }
    }
    if (rotations && !separable) {
        for (int i = 0; i < n_comp; i++) {
            const double *Rm = rotations + 9 * (size_t)i;
            if (std::isnan(Rm[0])) continue;
            bool identity = true, finite = true;
            for (int k = 0; k < 9; k++) {
                finite = finite && std::isfinite(Rm[k]);
                identity = identity && (Rm[k] == ((k % 4 == 0) ? 1.0 : 0.0));
            }
            if (!finite) return fail(ctx, CW_E_INVALID, "cw_set_potential_general: non-finite rotation matrix");
            if (identity) continue;
            // orthonormal rows, or grad = R^T grad' is not the gradient
            for (int r = 0; r < 3; r++)
                for (int c = r; c < 3; c++) {
                    const double d = Rm[3 * r] * Rm[3 * c] + Rm[3 * r + 1] * Rm[3 * c + 1] + Rm[3 * r + 2] * Rm[3 * c + 2];
                    if (fabs(d - (r == c ? 1.0 : 0.0)) > 1e-9)
                        return fail(ctx, CW_E_INVALID, "cw_set_potential_general: rotation matrix is not orthogonal");
                }
            P.has_R[i] = 1;
            for (int k = 0; k < 9; k++) P.R[i][k] = Rm[k];
            general = true;
        }
    }
    for (int i = 0; i < n_comp; i++) {
        const int t = types[i];
        if (t < CW_COMP_MIYAMOTO_NAGAI || t > CW_COMP_POWERLAW_CUTOFF)
            return fail(ctx, CW_E_UNSUPPORTED, "cw_set_potential: unknown component type");
        if (t >= CW_COMP_LONG_MURALI_BAR) general = true;
        P.type[i] = t;
        P.GM[i] = G * params[3 * i + 0];
        P.p1[i] = params[3 * i + 1];
        P.p2[i] = (t == CW_COMP_NFW_SPHERICAL) ? 1.0 / params[3 * i + 1] : params[3 * i + 2];
        if (t == CW_COMP_PLUMMER || t == CW_COMP_ISOCHRONE) P.p2[i] = params[3 * i + 1] * params[3 * i + 1];  // b^2
        if (t == CW_COMP_NFW_TRIAXIAL) {
            if (stride < 5) return fail(ctx, CW_E_INVALID, "cw_set_potential_ex: a triaxial NFW needs 5 parameters (m, r_s, a, b, c)");
            const double *pp = params_in + (size_t)stride * i;
            if (!(pp[2] > 0.0 && pp[3] > 0.0 && pp[4] > 0.0)) return fail(ctx, CW_E_INVALID, "cw_set_potential_ex: a, b, c must be > 0");
            P.p2[i] = 1.0 / pp[1];                       // 1 / r_s
            P.q2[i][0] = 1.0 / (pp[2] * pp[2]);          // 1 / a^2, 1 / b^2, 1 / c^2
            P.q2[i][1] = 1.0 / (pp[3] * pp[3]);
            P.q2[i][2] = 1.0 / (pp[4] * pp[4]);
        }
        if (t == CW_COMP_LOG_TRIAXIAL) {
            if (stride < 5) return fail(ctx, CW_E_INVALID, "cw_set_potential_ex: a triaxial logarithmic halo needs 5 parameters");
            const double *pp = params_in + (size_t)stride * i;
            if (!(pp[2] > 0.0 && pp[3] > 0.0 && pp[4] > 0.0)) return fail(ctx, CW_E_INVALID, "cw_set_potential_ex: q1, q2, q3 must be > 0");
            P.GM[i] = pp[0] * pp[0];                     // v_c^2
            P.p1[i] = pp[1] * pp[1];                     // r_h^2
            P.q2[i][0] = 1.0 / (pp[2] * pp[2]);
            P.q2[i][1] = 1.0 / (pp[3] * pp[3]);
            P.q2[i][2] = 1.0 / (pp[4] * pp[4]);
        }
        if (t == CW_COMP_LONG_MURALI_BAR) {
            if (stride < 4) return fail(ctx, CW_E_INVALID, "cw_set_potential: a Long-Murali bar needs 4 parameters (m, a, b, c)");
            const double *pp = params_in + (size_t)stride * i;
            if (!(pp[1] > 0.0 && pp[2] >= 0.0 && pp[3] >= 0.0 && pp[2] + pp[3] > 0.0))
                return fail(ctx, CW_E_INVALID, "cw_set_potential: Long-Murali bar needs a > 0, b, c >= 0, b + c > 0");
            P.pe[i][0] = pp[1]; P.pe[i][1] = pp[2]; P.pe[i][2] = pp[3] * pp[3];
        }
        if (t == CW_COMP_LEESUTO_NFW) {
            if (stride < 5) return fail(ctx, CW_E_INVALID, "cw_set_potential: a Lee-Suto NFW needs 5 parameters (v_c, r_s, a, b, c)");
            const double *pp = params_in + (size_t)stride * i;
            if (!(pp[1] > 0.0 && pp[2] > 0.0 && pp[3] > 0.0 && pp[4] > 0.0))
                return fail(ctx, CW_E_INVALID, "cw_set_potential: Lee-Suto NFW needs r_s, a, b, c > 0");
            const double eb2 = 1.0 - (pp[3] / pp[2]) * (pp[3] / pp[2]), ec2 = 1.0 - (pp[4] / pp[2]) * (pp[4] / pp[2]);
            const double ln2 = 0.69314718055994530942;
            P.GM[i] = pp[0] * pp[0] / (ln2 - 0.5 + (ln2 - 0.75) * eb2 + (ln2 - 0.75) * ec2);   // phi0
            P.pe[i][0] = 1.0 / pp[1]; P.pe[i][1] = eb2; P.pe[i][2] = ec2;
        }
        if (t == CW_COMP_POWERLAW_CUTOFF) {
            const double sg = 1.5 - 0.5 * params[3 * i + 1];
            if (!(sg > 0.0) || !(params[3 * i + 2] > 0.0))
                return fail(ctx, CW_E_INVALID, "cw_set_potential: power-law cutoff needs alpha < 3 and r_c > 0");
            P.pe[i][0] = sg;
            P.pe[i][1] = 1.0 / (params[3 * i + 2] * params[3 * i + 2]);
            P.pe[i][2] = lgamma(sg);
            P.pe[i][3] = (sg > 0.5) ? lgamma(sg - 0.5) : NAN;     // the value needs alpha < 2; the gradient does not
        }
        if (t == CW_COMP_BURKERT) {
            P.GM[i] = 3.14159265358979323846 * G * params[3 * i + 0] * params[3 * i + 1];   // pi G rho r0
            P.p2[i] = 1.0 / params[3 * i + 1];                                               // 1 / r0
        }
        if (t == CW_COMP_LOGARITHMIC) {
            if (!(params[3 * i + 2] > 0.0)) return fail(ctx, CW_E_INVALID, "cw_set_potential: logarithmic q3 must be > 0");
            P.GM[i] = params[3 * i + 0] * params[3 * i + 0];                 // v_c^2
            P.p1[i] = params[3 * i + 1] * params[3 * i + 1];                 // r_h^2
            P.p2[i] = 1.0 / (params[3 * i + 2] * params[3 * i + 2]);         // 1 / q3^2
        }
    }
    // recognise the two Milky Way shapes that have specialised kernels
    auto is = [&](int i, int t) { return i < n_comp && types[i] == t; };
    P.kind = general ? POT_GENERAL : POT_GENERIC;
    int nmn = 0;
    while (nmn < n_comp && types[nmn] == CW_COMP_MIYAMOTO_NAGAI) nmn++;
    if (!general && (nmn == 1 || nmn == 3) && n_comp == nmn + 3 && is(nmn, CW_COMP_HERNQUIST) &&
        is(nmn + 1, CW_COMP_HERNQUIST) && is(nmn + 2, CW_COMP_NFW_SPHERICAL)) {
        bool shared_b = true;
        for (int i = 1; i < nmn; i++) shared_b = shared_b && (params[3 * i + 2] == params[2]);
        if (nmn == 1) P.kind = POT_MW_V1;
        else if (shared_b) P.kind = POT_MW_V2;
        if (P.kind != POT_GENERIC) {
            for (int i = 0; i < nmn; i++) {
                P.mn_GM[i] = P.GM[i];
                P.mn_a[i] = params[3 * i + 1];
                P.mn_GMa[i] = P.GM[i] * params[3 * i + 1];
                P.mn_b2[i] = params[3 * i + 2] * params[3 * i + 2];
            }
            for (int k = 0; k < 2; k++) {
                P.h_GM[k] = P.GM[nmn + k];
                P.h_c[k] = params[3 * (nmn + k) + 1];
            }
            P.nfw_GM = P.GM[nmn + 2];
            P.nfw_rs = params[3 * (nmn + 2) + 1];
            P.nfw_inv_rs = 1.0 / P.nfw_rs;
        }
    }
    if (P.kind == POT_GENERAL) {
        // static and diagonal throughout (only rotated frames): the lighter kernels of POT_ROTATED, same bits
        bool light = true;
        for (int i = 0; i < n_comp; i++)
            light = light && P.k_off[i + 1] == P.k_off[i] && types[i] < CW_COMP_LONG_MURALI_BAR;
        if (light && !getenv("CW_NO_ROTATED_KIND")) P.kind = POT_ROTATED;
        // gala's LM10Potential: the component list the kernels of POT_LM10 are compiled for
        if (P.kind == POT_ROTATED && n_comp == 3 && types[0] == CW_COMP_MIYAMOTO_NAGAI && types[1] == CW_COMP_HERNQUIST &&
            types[2] == CW_COMP_LOG_TRIAXIAL && !P.has_R[0] && !P.has_R[1] && P.has_R[2] && !getenv("CW_NO_LM10_KIND"))
            P.kind = POT_LM10;
    }
    if (P.kind == POT_GENERAL || P.kind == POT_ROTATED || P.kind == POT_LM10)
        for (int i = n_comp - 1; i >= 0; i--) {
            const bool plain = !P.has_R[i] && P.k_off[i + 1] == P.k_off[i] && types[i] < CW_COMP_LONG_MURALI_BAR;
            P.run_end[i] = !plain ? 0 : (i + 1 < n_comp && P.run_end[i + 1] > i + 1) ? P.run_end[i + 1] : i + 1;
        }
    if (separable) P.kind = (P.kind == POT_MW_V2) ? POT_MW_V2_T : POT_GENERIC_T;
    pot_fill_constants(P);
    ctx->pot = P;
    ctx->has_pot = true;
    return CW_OK;
}

// ---------------------------------------------------------------------------------------------
// cw_integrate / cw_count_nodes
//
// CW_MEM_DEVICE: one shard on the context's first device, every pointer used as it is.
// CW_MEM_HOST:   the batch is cut by index over the device slots (one host thread per slot) and, within a slot, into
//                chunks that rotate over N_LANES streams, so the H2D copy of chunk c+1 and the D2H copy of chunk c-1
//                overlap the kernels of chunk c.  Caller arrays that are page-locked are DMA-ed directly; pageable ones
//                (plain numpy) go through the lane's page-locked staging buffers, copied by the host thread pool.
// ---------------------------------------------------------------------------------------------

namespace {

constexpr int64_t SORT_THRESHOLD = 4096;
// About 8 chunks per device, between 2^16 orbits (below that launch overheads show) and 2^20 (bounds the exposed first
// copy).
constexpr int64_t HOST_CHUNK_MAX = 1 << 20, HOST_CHUNK_MIN = 1 << 16;
constexpr int64_t TRAJ_SHARD_COLS = (int64_t)1 << 26;          // trajectory columns per host-mode chunk (3.2 GB)
constexpr int64_t TRAJ_SHARD_COLS_STAGED = (int64_t)1 << 23;   // ... when they pass through the staging ring (400 MB)
static int64_t traj_shard_cols(bool staged)
{
    const char *e = getenv("CW_TRAJ_SHARD_COLS");       // test hook: force many small shards
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (int64_t)v : (staged ? TRAJ_SHARD_COLS_STAGED : TRAJ_SHARD_COLS);
}
// Q > 0: the number of orbits one full grid of the kernel holds (persistent lanes).  Chunks are whole multiples of it,
// so no chunk ends with a wave that fills a fraction of the GPU (10^6 orbits in LM10 at Q = 113 664: 127k-orbit
// chunks = 1.12 waves each cost a third of the end-to-end throughput).
static int64_t host_chunk(int64_t n_on_device, int64_t Q, bool staged)
{
    int64_t c = (n_on_device + 7) / 8;
    // Pageable caller arrays pass through the lanes' page-locked staging buffers, which are as large as the largest
    // chunk and cost 0.45 ms per MB to pin when a context is first used: two waves per chunk (24 + 18 MB per lane)
    // instead of up to six keep that below 70 ms for a caller that runs one call.
    if (staged && Q > 0) c = std::min(c, 2 * Q);
    if (const char *e = getenv("CW_HOST_CHUNK")) {      // experiment hook: fixed chunk size in orbits
        const long long v = atoll(e);
        if (v > 0) return (int64_t)v;
    }
    if (Q <= 0 || getenv("CW_NO_WAVE_CHUNKS")) {
        c = (c + 4095) & ~int64_t(4095);
        return std::min(HOST_CHUNK_MAX, std::max(HOST_CHUNK_MIN, c));
    }
    c = ((c + Q - 1) / Q) * Q;
    const int64_t cmax = std::max<int64_t>(Q, (HOST_CHUNK_MAX / Q) * Q);
    return std::min(cmax, std::max(Q, c));
}

// orbit counts of the chunks of one device slot holding M orbits
static std::vector<int64_t> chunk_sizes(int64_t M, bool pipelined, int64_t Q, bool staged)
{
    std::vector<int64_t> sizes;
    const int64_t chunk = pipelined ? host_chunk(M, Q, staged) : M;
    // The first H2D copy and the last D2H copy have nothing to hide behind: ramp the chunk size up from a fraction
    // of it at the start and back down at the end (the kernels in between keep the full size).
    if (!pipelined || M < 4 * chunk) {
        for (int64_t c = 0; c < M; c += chunk) sizes.push_back(std::min(chunk, M - c));
        return sizes;
    }
    std::vector<int64_t> head, tail;
    int64_t used = 0;
    if (Q > 0) {
        // Q/2, Q, 2Q, ...: the first H2D and the last D2H move half a wave (6 MB); the fractional waves do not idle
        // the GPU, the next chunk's kernel (another stream) fills the SMs they leave free.  Measured at 1.25e6 orbits
        // (one rank of the 8-GPU run): 8.09 ms starting at Q, 8.03 at Q/2, 8.55 at Q/4, 8.46 at Q/8.
        int64_t q0 = Q / 2;
        if (const char *e = getenv("CW_HOST_RAMP_DIV")) { const int v = atoi(e); if (v >= 1) q0 = Q / v; }
        q0 = std::max<int64_t>(q0, 4096);
        for (int64_t sz = q0; sz < chunk && used + 2 * sz < M; sz *= 2) { head.push_back(sz); used += sz; }
        for (int64_t sz = q0; sz < chunk && used + 2 * sz < M; sz *= 2) { tail.push_back(sz); used += sz; }
    } else {
        for (int64_t sz = std::max<int64_t>(chunk / 8, HOST_CHUNK_MIN / 2); sz < chunk; sz *= 2) { head.push_back(sz); used += sz; }
        for (int64_t sz = std::max<int64_t>(chunk / 4, HOST_CHUNK_MIN / 2); sz < chunk; sz *= 2) { tail.push_back(sz); used += sz; }
    }
    int64_t mid = M - used;
    sizes = head;
    const size_t first_mid = sizes.size();
    while (mid > 0) { const int64_t sz = std::min(chunk, mid); sizes.push_back(sz); mid -= sz; }
    // a short remainder does not get a launch of its own (a quarter-wave chunk occupies the GPU for a whole wave's
    // time: 1.0 ms for 0.25 waves measured); the previous chunk takes it, its last partial wave then overlaps the
    // next chunk's first one on another stream
    if (sizes.size() > first_mid + 1 && sizes.back() < chunk / 2) {
        const int64_t r = sizes.back();
        sizes.pop_back();
        sizes.back() += r;
    }
    for (size_t k = tail.size(); k-- > 0;) sizes.push_back(tail[k]);
    return sizes;
}

// size of the CUB temp storage for sorting m (key, value) pairs
size_t sort_temp_bytes(int64_t m)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                              (const int32_t *)nullptr, (int32_t *)nullptr, (int)m);
    return bytes;
}

struct Plan {
    IntegArgs A;
    int32_t *iota = nullptr, *keys_out = nullptr, *keys_in = nullptr, *order = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_bytes = 0;
    bool sort = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    bool timed = false;
};

// failures inside a slot's worker are reported through the slot, not the context (several slots run at once)
#define CW_SCK(slot, call)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (slot).error = _b;                                                                    \
            return CW_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

static int slot_fail(Slot &s, int code, const char *msg)
{
    s.error = msg;
    return code;
}

// grid pre-pass, queue ordering and the integrator kernel of one shard / chunk, on pl.stream
static int run_kernels(cw_ctx *ctx, Slot &s, const cw_opts *o, Plan &pl, bool integrate, bool store_all)
{
    cudaStream_t stream = pl.stream;
    LaunchCfg cfg{s.sm_count, stream};
    IntegArgs &A = pl.A;
    CW_SCK(s, cudaMemsetAsync(A.queue, 0, 64, stream));
    CW_SCK(s, cudaEventRecord(pl.ev_t0, stream));
    CW_SCK(s, launch_grid_prepass(A, pl.sort ? pl.iota : nullptr, pl.sort ? pl.keys_in : nullptr, cfg));
    s.launches += 1;
    if (integrate) {
        if (pl.sort) {
            // Longest-processing-time-first queue.  Leapfrog: cost = node count.  DOP853: cost = node
            // count x step attempts per interval, measured by a short pilot integration (the persistent
            // short-period orbits near the nucleus need 10-50x the attempts of a disc orbit).
            const int32_t *keys = pl.keys_in;      // node count << 4 | trajectory column phase
            if (o->integrator != CW_LEAPFROG) {
                A.cost_key = pl.keys_in;
                A.pilot_intervals = 16;
                A.order = nullptr;
                CW_SCK(s, launch_dop853_pilot(ctx->pot, A, cfg));
                s.launches += 1;
                CW_SCK(s, cudaMemsetAsync(A.queue, 0, 8, stream));
                A.order = pl.order;
                keys = pl.keys_in;
            }
            CW_SCK(s, cub::DeviceRadixSort::SortPairsDescending(pl.sort_tmp, pl.sort_bytes, keys,
                                                                pl.keys_out, (const int32_t *)pl.iota, A.order,
                                                                (int)A.n, 0, 32, stream));
        }
        CW_SCK(s, cudaEventRecord(pl.ev_k0, stream));
        if (o->integrator == CW_LEAPFROG) CW_SCK(s, launch_leapfrog(ctx->pot, A, store_all, cfg));
        else CW_SCK(s, launch_dop853(ctx->pot, A, store_all, o->integrator == CW_DOP853_DENSE, cfg));
        s.launches += 1;
        CW_SCK(s, cudaEventRecord(pl.ev_k1, stream));
        pl.timed = true;
        if (store_all && A.traj_t) {
            CW_SCK(s, launch_traj_times(A, cfg));
            s.launches += 1;
        }
    }
    return CW_OK;
}

static int validate(cw_ctx *ctx, const cw_opts *o, const cw_batch *B, bool integrate)
{
    if (!ctx || !o || !B) return CW_E_INVALID;
    if (B->n_orbits < 0 || B->n_orbits > (int64_t)INT32_MAX) return fail(ctx, CW_E_INVALID, "n_orbits out of range (at most 2^31 - 1 per call)");
    if (o->mem_space != CW_MEM_HOST && o->mem_space != CW_MEM_DEVICE) return fail(ctx, CW_E_INVALID, "bad mem_space");
    const bool compact = B->n_w0 != 0 || B->w0_index || B->grid_class || B->classes || B->ev_off32;
    if (compact && o->mem_space != CW_MEM_HOST)
        return fail(ctx, CW_E_INVALID, "n_w0 / w0_index / grid_class / ev_off32 are CW_MEM_HOST forms");
    if (B->n_orbits > 0) {
        if (!B->t1 || (integrate && !B->w0)) return fail(ctx, CW_E_INVALID, "null input array");
        if (B->grid_class) {
            if (!B->classes || B->n_classes < 1 || B->n_classes > 256) return fail(ctx, CW_E_INVALID, "grid_class needs 1..256 classes");
        } else if (!B->t2 || !B->dt || !B->to_myr || !B->grid_flags) {
            return fail(ctx, CW_E_INVALID, "null input array");
        }
        if (B->n_w0 < 0 || B->n_w0 > B->n_orbits) return fail(ctx, CW_E_INVALID, "n_w0 out of range");
        if (B->n_w0 > 0 && B->n_w0 < B->n_orbits && !B->w0_index) return fail(ctx, CW_E_INVALID, "n_w0 < n_orbits needs w0_index");
        if (B->ev_off && B->ev_off32) return fail(ctx, CW_E_INVALID, "give ev_off or ev_off32, not both");
    }
    if (o->integrator != CW_LEAPFROG && o->integrator != CW_DOP853 && o->integrator != CW_DOP853_DENSE) return fail(ctx, CW_E_UNSUPPORTED, "unknown integrator");
    if (ctx->slots.empty()) return fail(ctx, CW_E_NO_DEVICE, "context has no device");
    return CW_OK;
}

// =============================================================================================
// CW_MEM_DEVICE
// =============================================================================================
static int integrate_device(cw_ctx *ctx, const cw_opts *o, const cw_batch *B, const cw_result *R, bool integrate,
                            int32_t *n_nodes_out, int32_t *ev_step_out)
{
    const int64_t n = B->n_orbits;
    const bool store_all = integrate && R->traj_off;
    DeviceGuard guard;
    Slot &s = ctx->slots[0];
    s.error.clear();
    s.launches = 0;
    CW_CK(ctx, cudaSetDevice(s.dev));
    int64_t K = 0, Lt = 0;
    if (B->ev_off) CW_CK(ctx, cudaMemcpy(&K, B->ev_off + n, 8, cudaMemcpyDeviceToHost));
    if (store_all) CW_CK(ctx, cudaMemcpy(&Lt, R->traj_off + n, 8, cudaMemcpyDeviceToHost));
    if (K < 0 || Lt < 0) return fail(ctx, CW_E_INVALID, "offsets must be non-decreasing");
    if (K > 0 && (!B->ev_time || !B->ev_dv)) return fail(ctx, CW_E_INVALID, "ev_off lists events but ev_time / ev_dv is null");

    Plan pl;
    pl.stream = o->stream ? (cudaStream_t)o->stream : s.lanes[0];
    pl.sort = integrate && n >= SORT_THRESHOLD;
    pl.sort_bytes = pl.sort ? sort_temp_bytes(n) : 0;
    size_t need = 0;
    auto add = [&](size_t count, size_t elt) { need += Arena::padded(count, elt); };
    add(K + 1, 8);
    add(n, 8);                                  // t_last
    if (pl.sort) { add(n, 4); add(n, 4); add(n, 4); add(n, 4); add(pl.sort_bytes, 1); }
    add(n, 4); add(K + 1, 4); add(n, 4);        // n_nodes / ev_step / status if the caller passed NULL
    need += 4096;
    // The scratch arena is the context's: a previous no_sync call may still be running on another stream.  Order
    // this call behind it on the device (cheap), and wait on the host only when the buffers have to move.
    if (s.has_pending) {
        if (need > s.lane_arena[0].cap) CW_CK(ctx, cudaEventSynchronize(s.pending));
        else CW_CK(ctx, cudaStreamWaitEvent(pl.stream, s.pending, 0));
        s.has_pending = false;
    }
    if (need > s.lane_arena[0].cap) cudaStreamSynchronize(s.lanes[0]);
    Arena &ar = s.lane_arena[0];
    CW_CK(ctx, ar.reserve(need));
    CW_CK(ctx, s.meta.reserve(4096));
    CW_CK(ctx, s.ensure_events(3));
    pl.ev_t0 = s.events[0]; pl.ev_k0 = s.events[1]; pl.ev_k1 = s.events[2];

    IntegArgs &A = pl.A;
    memset(&A, 0, sizeof A);
    A.n = n;
    A.atol = o->atol > 0 ? o->atol : 1e-10;
    A.rtol = o->rtol > 0 ? o->rtol : 1e-10;
    A.nmax = o->nmax;
    A.ev_tk = ar.take<double>(K + 1);
    if (integrate && o->integrator == CW_DOP853_DENSE) A.t_last = ar.take<double>(n);
    unsigned long long *qs = s.meta.take<unsigned long long>(8);
    A.queue = qs;
    A.stats = qs + 1;
    if (pl.sort) {
        pl.iota = ar.take<int32_t>(n);
        pl.keys_out = ar.take<int32_t>(n);
        pl.keys_in = ar.take<int32_t>(n);
        pl.order = ar.take<int32_t>(n);
        A.order = pl.order;
        pl.sort_tmp = ar.take<char>(pl.sort_bytes);
    }
    A.w0 = B->w0; A.t1 = B->t1; A.t2 = B->t2; A.dt = B->dt; A.to_myr = B->to_myr; A.flags = B->grid_flags;
    if (B->ev_off && K > 0) { A.ev_off = B->ev_off; A.ev_base = 0; A.ev_time = B->ev_time; A.ev_dv = B->ev_dv; }
    int32_t *nn = ar.take<int32_t>(n), *es = ar.take<int32_t>(K + 1), *st = ar.take<int32_t>(n);
    int32_t *nn_user = integrate ? R->n_nodes : n_nodes_out, *es_user = integrate ? R->ev_step : ev_step_out;
    A.n_nodes = nn_user ? nn_user : nn;
    A.ev_step = es_user ? es_user : es;
    A.status = (integrate && R->status) ? R->status : st;
    A.w_final = integrate ? R->w_final : nullptr;
    if (store_all) {
        A.traj_off = R->traj_off; A.traj_base = 0; A.traj_total = Lt;
        A.traj_pos = R->traj_pos; A.traj_vel = R->traj_vel; A.traj_t = R->traj_t;
    }
    if (!A.ev_tk || !A.queue || !A.n_nodes || !A.ev_step || !A.status)
        return fail(ctx, CW_E_CUDA, "device arena exhausted (internal sizing error)");

    int rc = run_kernels(ctx, s, o, pl, integrate, store_all);
    ctx->last_launches = s.launches;
    if (rc != CW_OK) { cudaStreamSynchronize(pl.stream); ctx->last_error = s.error; return rc; }
    if (integrate && R->stats)
        CW_CK(ctx, cudaMemcpyAsync(R->stats, A.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, pl.stream));
    if (integrate && R->n_failed)
        CW_CK(ctx, cudaMemcpyAsync(R->n_failed, A.stats + 4, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, pl.stream));
    if (o->no_sync) {
        CW_CK(ctx, cudaEventRecord(s.pending, pl.stream));
        s.has_pending = true;
        return CW_OK;
    }
    CW_CK(ctx, cudaStreamSynchronize(pl.stream));
    {
        unsigned long long mx = 0;
        if (integrate && cudaMemcpy(&mx, A.stats + 5, 8, cudaMemcpyDeviceToHost) == cudaSuccess) ctx->last_max_attempts = (int64_t)mx;
    }
    if (pl.timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pl.ev_k0, pl.ev_k1) == cudaSuccess) s.last_kernel_ms += ms;
        if (cudaEventElapsedTime(&ms, pl.ev_t0, pl.ev_k1) == cudaSuccess) s.last_total_ms += ms;
    }
    return CW_OK;
}

// =============================================================================================
// CW_MEM_HOST
// =============================================================================================

// expands the compact host forms (cw_batch ABI 2) of one chunk into the plain layout the kernels read
struct UnpackArgs {
    int64_t m1, m2, ua;
    const double *u_w0, *u_t1;
    const int32_t *u_idx;
    const uint8_t *u_cls;
    const cw_grid_class *u_ctab;
    int32_t n_classes;
    const void *u_off1, *u_off2;
    int32_t off_is32, do_w0;
    const int64_t *u_toff1, *u_toff2;
    double *k_w0, *k_t1, *k_t2, *k_dt, *k_tm;
    int32_t *k_fl;
    int64_t *k_evoff, *k_toff;
};

__global__ void __launch_bounds__(256) unpack_chunk_kernel(const __grid_constant__ UnpackArgs U)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = U.m1 + U.m2;
    if (i > m) return;
    if (i < m) {
        if (U.do_w0 || U.k_t1) {
            int64_t j = i;
            bool bad = false;
            if (i >= U.m1) {
                j = (int64_t)U.u_idx[i - U.m1] - U.ua;
                bad = (j < 0 || j >= U.m1);        // not this chunk's source: the caller's w0_index is not sorted
                if (bad) j = 0;
            }
            if (U.do_w0)
                for (int c = 0; c < 6; c++) U.k_w0[c * m + i] = U.u_w0[c * U.m1 + j];
            // a NaN start time makes the pre-pass reject the orbit (status CW_ORBIT_INCONSISTENT)
            if (U.k_t1) U.k_t1[i] = bad ? __longlong_as_double(0x7ff8000000000000LL) : U.u_t1[j];
        }
        if (U.u_cls) {
            const int c = U.u_cls[i];
            const bool ok = c < U.n_classes;
            const cw_grid_class g = U.u_ctab[ok ? c : 0];
            U.k_t2[i] = g.t2;
            U.k_dt[i] = ok ? g.dt : 0.0;           // dt = 0 is rejected by the pre-pass
            U.k_tm[i] = g.to_myr;
            U.k_fl[i] = g.grid_flags;
        }
    }
    auto off_at = [&](const void *p, int64_t k) -> int64_t {
        return U.off_is32 ? (int64_t)((const int32_t *)p)[k] : ((const int64_t *)p)[k];
    };
    if (U.k_evoff) {
        int64_t v;
        if (i <= U.m1) v = off_at(U.u_off1, i) - off_at(U.u_off1, 0);
        else v = (off_at(U.u_off1, U.m1) - off_at(U.u_off1, 0)) + off_at(U.u_off2, i - U.m1) - off_at(U.u_off2, 0);
        U.k_evoff[i] = v;
    }
    if (U.k_toff) {
        int64_t v;
        if (i <= U.m1) v = U.u_toff1[i] - U.u_toff1[0];
        else v = (U.u_toff1[U.m1] - U.u_toff1[0]) + U.u_toff2[i - U.m1] - U.u_toff2[0];
        U.k_toff[i] = v;
    }
}

struct Seg {
    size_t off = 0, bytes = 0;
};

// One unit of host-mode work: the sources [ua, ub) with the orbits that start from them -- the primaries [ua, ub) and
// the secondaries n_src + [sa, sb) -- on one stream lane of one device slot.
struct Chunk {
    int slot = 0, lane = 0;
    int64_t ua = 0, ub = 0, sa = 0, sb = 0, m1 = 0, m2 = 0, m = 0;
    int64_t e1a = 0, K1 = 0, e2a = 0, K2 = 0, K = 0;      // event rows of the two orbit ranges
    int64_t c1a = 0, L1 = 0, c2a = 0, L2 = 0, Lt = 0, Ltp = 0;   // trajectory columns
    bool unpack = false, do_w0 = false;
    // upload region (mirrored by the lane's stage_in buffer)
    Seg u_w0, u_t1, u_idx, u_cls, u_ctab, u_t2, u_dt, u_tm, u_fl, u_off1, u_off2, u_evt, u_evd, u_toff1, u_toff2;
    size_t up_bytes = 0;
    // download region (mirrored by stage_out): results, then the trajectories
    size_t down_off = 0, down_bytes = 0;
    Seg d_wf, d_st, d_nn, d_es, d_tp, d_tt;
    // device only
    Seg k_w0, k_t1, k_grid, k_fl, k_evoff, k_toff, k_evtk, k_tlast, k_sort4, k_sorttmp;
    size_t total_bytes = 0;
    size_t sort_bytes = 0;
    bool sort = false;
    bool staged_in = false, staged_out = false;
    // drain work: stage_out -> caller arrays, once d2h_done has fired
    std::vector<CopyTask> drain;
    cudaEvent_t d2h_done = nullptr;
    bool drained = true;
    Plan plan;
};

// resolved view of the caller's batch for a host-mode call
struct HostCall {
    cw_ctx *ctx = nullptr;
    const cw_opts *o = nullptr;
    const cw_batch *B = nullptr;
    cw_result R;                      // outputs (n_nodes / ev_step of cw_count_nodes included)
    bool integrate = false, store_all = false;
    int64_t n = 0, n_src = 0, n_sec = 0;
    const int32_t *idx = nullptr;
    bool has_ev = false, want_ev = false, classes = false, off32 = false;
    int64_t total_cols = 0;
    // which caller arrays can be DMA-ed as they are (page-locked / registered); the others are staged
    bool p_w0, p_t1, p_t2, p_dt, p_tm, p_fl, p_cls, p_ctab, p_idx, p_off, p_evt, p_evd, p_toff;
    bool p_wf, p_st, p_nn, p_es, p_tp, p_tv, p_tt;

    // CW_DEBUG: host wall clock of the call (every slot thread prints its own milestones relative to t_call)
    bool trace = false;
    std::chrono::steady_clock::time_point t_call;
    double ms_since_call() const
    {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
    }

    int64_t ev_at(int64_t i) const { return off32 ? (int64_t)B->ev_off32[i] : B->ev_off[i]; }
    // number of secondaries whose source is < u
    int64_t lb(int64_t u) const
    {
        if (!idx) return 0;
        const int32_t key = (int32_t)std::min<int64_t>(u, INT32_MAX);
        return (int64_t)(std::lower_bound(idx, idx + n_sec, key) - idx);
    }
    int64_t orbits_before(int64_t u) const { return u + lb(u); }
    int64_t cols_before(int64_t u) const
    {
        return (R.traj_off[u] - R.traj_off[0]) + (R.traj_off[n_src + lb(u)] - R.traj_off[n_src]);
    }
};

static inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// byte layout of a chunk in its lane's arena
static void layout_chunk(const HostCall &H, Chunk &c)
{
    size_t cur = 0;
    auto seg = [&](size_t bytes) { Seg s; s.off = cur; s.bytes = bytes; cur += pad256(bytes); return s; };
    const size_t m = (size_t)c.m, m1 = (size_t)c.m1, m2 = (size_t)c.m2, K = (size_t)c.K;
    const size_t osz = H.off32 ? 4 : 8;
    c.unpack = (c.m2 > 0) || H.classes || (H.off32 && H.want_ev);
    c.do_w0 = H.integrate && c.m2 > 0;
    // ---- upload
    if (H.integrate) c.u_w0 = seg(6 * m1 * 8);
    c.u_t1 = seg(m1 * 8);
    if (m2) c.u_idx = seg(m2 * 4);
    if (H.classes) {
        c.u_cls = seg(m);
        c.u_ctab = seg((size_t)H.B->n_classes * sizeof(cw_grid_class));
    } else {
        c.u_t2 = seg(m * 8); c.u_dt = seg(m * 8); c.u_tm = seg(m * 8); c.u_fl = seg(m * 4);
    }
    if (H.want_ev && K > 0) {
        c.u_off1 = seg((m1 + 1) * osz);
        if (m2) c.u_off2 = seg((m2 + 1) * osz);
        c.u_evt = seg(K * 8);
        if (H.integrate) c.u_evd = seg(3 * K * 8);
    }
    if (H.store_all) {
        c.u_toff1 = seg((m1 + 1) * 8);
        if (m2) c.u_toff2 = seg((m2 + 1) * 8);
    }
    c.up_bytes = cur;
    // ---- download
    c.down_off = cur;
    if (H.integrate) c.d_wf = seg(6 * m * 8);
    c.d_st = seg(m * 4);
    c.d_nn = seg(m * 4);
    c.d_es = seg((K + 1) * 4);
    if (H.store_all) {
        c.d_tp = seg(6 * (size_t)c.Ltp * 8 + 256);
        c.d_tt = seg(((size_t)c.Lt + 1) * 8);
    }
    c.down_bytes = cur - c.down_off;
    // ---- device only
    if (c.unpack) {
        if (c.do_w0) { c.k_w0 = seg(6 * m * 8); c.k_t1 = seg(m * 8); }
        if (H.classes) { c.k_grid = seg(3 * m * 8); c.k_fl = seg(m * 4); }
        if (H.want_ev && K > 0) c.k_evoff = seg((m + 1) * 8);
        if (H.store_all) c.k_toff = seg((m + 1) * 8);
    }
    c.k_evtk = seg((K + 1) * 8);
    if (H.integrate && H.o->integrator == CW_DOP853_DENSE) c.k_tlast = seg(m * 8);
    c.sort = H.integrate && c.m >= SORT_THRESHOLD;
    c.sort_bytes = c.sort ? sort_temp_bytes(c.m) : 0;
    if (c.sort) { c.k_sort4 = seg(4 * pad256(m * 4)); c.k_sorttmp = seg(c.sort_bytes); }
    c.total_bytes = cur + 1024;
}

// cut the batch: by orbit count over the device slots, then into pipelined chunks
static int make_chunks(HostCall &H, std::vector<Chunk> &chunks)
{
    cw_ctx *ctx = H.ctx;
    const int64_t n = H.n;
    const int nslots = (int)std::min<int64_t>((int64_t)ctx->slots.size(), std::max<int64_t>(1, n));
    // first unit u >= lo with orbits_before(u) >= target
    auto unit_for = [&](int64_t target, int64_t lo, int64_t hi) {
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (H.orbits_before(mid) >= target) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    const bool traj_staged = H.store_all && !(H.p_tp && H.p_tv && (!H.R.traj_t || H.p_tt));
    const int64_t col_lim = traj_shard_cols(traj_staged);
    // ---- the cut over the device slots.  Slices stay contiguous (no permutation of the caller's arrays, no gather on
    // the host), but they hold equal WORK, not equal orbit counts: the grid steps of every 4096th source unit are
    // read (a thousand-odd cache misses: 0.05 ms), taken for its block, and the cut points are where the running sum
    // passes k / nslots of the total.  Secondaries are taken to be spread evenly over the sources.  A population in
    // random order is cut as before; one sorted by age (steps grow along the index) no longer leaves the last device
    // with twice its share.
    std::vector<int64_t> slot_end(nslots, H.n_src);
    const cw_batch *B = H.B;
    if (nslots > 1) {
        const int64_t B0 = 4096;
        const int64_t nb = (H.n_src + B0 - 1) / B0;
        std::vector<double> cum(nb + 1, 0.0);
        for (int64_t b = 0; b < nb; b++) {
            const int64_t u = b * B0, ue = std::min(H.n_src, u + B0);
            double t2, dt;
            if (H.classes) {
                const int c = B->grid_class[u];
                const cw_grid_class g = B->classes[c < B->n_classes ? c : 0];
                t2 = g.t2; dt = g.dt;
            } else {
                t2 = B->t2[u]; dt = B->dt[u];
            }
            double steps = (dt != 0.0) ? (t2 - B->t1[u]) / dt : 1.0;
            if (!(steps >= 1.0)) steps = 1.0;                 // NaN, negative, empty grids
            if (steps > 1.0e6) steps = 1.0e6;
            cum[b + 1] = cum[b] + steps * (double)(ue - u);
        }
        int64_t b = 0;
        for (int d = 0; d + 1 < nslots; d++) {
            const double target = cum[nb] * (double)(d + 1) / (double)nslots;
            while (b < nb && cum[b + 1] <= target) b++;
            // the block that straddles the target is split in proportion (its orbits are taken to cost the same)
            int64_t u = b * B0;
            if (b < nb && cum[b + 1] > cum[b])
                u += (int64_t)((double)(std::min(H.n_src, u + B0) - u) * (target - cum[b]) / (cum[b + 1] - cum[b]));
            slot_end[d] = std::min(std::max(u, d > 0 ? slot_end[d - 1] : (int64_t)0), H.n_src);
        }
    }
    if (H.trace && nslots > 1)
        for (int d = 0; d < nslots; d++)
            fprintf(stderr, "[cw] slot %d takes source units [%lld, %lld)\n", d, (long long)(d ? slot_end[d - 1] : 0),
                    (long long)slot_end[d]);
    int64_t u_prev = 0;
    for (int d = 0; d < nslots; d++) {
        const int64_t u_lo = u_prev;
        const int64_t u_hi = slot_end[d];
        u_prev = u_hi;
        if (u_hi <= u_lo) continue;
        const int64_t base = H.orbits_before(u_lo);
        const int64_t M = H.orbits_before(u_hi) - base;
        // DOPRI853 runs are bound by their most expensive orbits, not by the copies (10^6 orbits: 2.4 ms of H2D
        // against 150 ms of kernel): cutting them into chunks would multiply the pilot launches and the tails
        // (measured 3.5e9 vs 6.2e9 intervals/s end to end), so they go to the device in one piece.
        const bool pipelined = !H.integrate || H.o->integrator == CW_LEAPFROG;
        int64_t Q = 0;
        bool one_chunk = false;
        if (pipelined && H.integrate && !H.store_all) {
            cudaSetDevice(ctx->slots[d].dev);
            Q = leapfrog_lane_capacity(ctx->pot, ctx->slots[d].sm_count);
            if (M < 4 * Q && !getenv("CW_HOST_CHUNK")) {
                // A batch of less than four waves whose copies are small beside its kernel time goes to the device in
                // ONE piece: nothing is gained by overlapping 0.5 ms of copies with 8 ms of kernel, and the queue is
                // then ordered longest-first over the whole batch instead of chunk by chunk (10^5 binaries with
                // Wagg2022 horizons, BASELINE config 2: 9.6 ms against 10.8 ms per call; 3 x 10^5: 25.3 against 29.2).
                // Kernel time from the grid steps of <= 256 sampled orbits at 1.8e11 steps/s, copies at 140 B per orbit
                // and 50 GB/s; short orbits (1000 steps: copies are half the kernel time) keep the pipelined cut.
                const int64_t stride = std::max<int64_t>(1, (u_hi - u_lo) / 256);
                double steps = 0.0;
                int64_t ns = 0;
                for (int64_t u = u_lo; u < u_hi; u += stride, ns++) {
                    double t2, dt;
                    if (H.classes) {
                        const int c = B->grid_class[u];
                        const cw_grid_class g = B->classes[c < B->n_classes ? c : 0];
                        t2 = g.t2; dt = g.dt;
                    } else {
                        t2 = B->t2[u]; dt = B->dt[u];
                    }
                    double st = (dt != 0.0) ? (t2 - B->t1[u]) / dt : 1.0;
                    if (!(st >= 1.0)) st = 1.0;
                    steps += std::min(st, 1.0e6);
                }
                const double kernel_s = (double)M * (steps / (double)std::max<int64_t>(ns, 1)) / 1.8e11;
                const double copy_s = (double)M * 140.0 / 50.0e9;
                one_chunk = copy_s < 0.15 * kernel_s;
            }
            // only batches of several waves: a batch of one or two (10^5 binaries with their own horizons, config 2)
            // pipelines better in the finer chunks than in wave-sized ones -- measured 1.34e11 against 0.90e11 steps/s
            if (getenv("CW_NO_WAVE_CHUNKS") || M < 4 * Q) Q = 0;
        }
        // (the per-orbit / per-event arrays decide; the few bytes of the class table may well be ordinary memory)
        const bool staged = !(H.p_w0 && H.p_t1 && H.p_t2 && H.p_dt && H.p_tm && H.p_fl && H.p_cls && H.p_idx &&
                              H.p_off && H.p_evt && H.p_evd && H.p_wf && H.p_st && H.p_nn && H.p_es);
        const std::vector<int64_t> sizes = one_chunk ? std::vector<int64_t>{M} : chunk_sizes(M, pipelined, Q, staged);
        int lane = 0;
        int64_t u = u_lo, cum = 0;
        for (size_t k = 0; k < sizes.size() && u < u_hi; k++) {
            cum += sizes[k];
            int64_t ue = (k + 1 == sizes.size()) ? u_hi : unit_for(base + cum, u + 1, u_hi);
            while (u < ue) {
                int64_t e1 = ue;
                if (H.store_all) {
                    // a chunk's trajectories live in its lane's buffer: also cut by trajectory COLUMNS; with the lane
                    // buffers re-used this bounds device memory however long the orbits are
                    const int64_t c0 = H.cols_before(u);
                    int64_t lo = u + 1, hi = ue;       // largest e in [u+1, ue] with cols <= limit (at least u+1)
                    while (lo < hi) {
                        const int64_t mid = lo + (hi - lo + 1) / 2;
                        if (H.cols_before(mid) - c0 <= col_lim) lo = mid; else hi = mid - 1;
                    }
                    e1 = lo;
                }
                Chunk c;
                c.slot = d;
                c.lane = lane++ % Slot::N_LANES;
                c.ua = u; c.ub = e1;
                c.sa = H.lb(u); c.sb = H.lb(e1);
                c.m1 = c.ub - c.ua; c.m2 = c.sb - c.sa; c.m = c.m1 + c.m2;
                if (H.has_ev) {
                    c.e1a = H.ev_at(c.ua); c.K1 = H.ev_at(c.ub) - c.e1a;
                    c.e2a = H.ev_at(H.n_src + c.sa); c.K2 = H.ev_at(H.n_src + c.sb) - c.e2a;
                    if (c.K1 < 0 || c.K2 < 0) return fail(ctx, CW_E_INVALID, "offsets must be non-decreasing");
                    c.K = H.want_ev ? c.K1 + c.K2 : 0;
                    if (!H.want_ev) c.K1 = c.K2 = 0;
                }
                if (H.store_all) {
                    const int64_t *to = H.R.traj_off;
                    c.c1a = to[c.ua]; c.L1 = to[c.ub] - c.c1a;
                    c.c2a = to[H.n_src + c.sa]; c.L2 = to[H.n_src + c.sb] - c.c2a;
                    if (c.L1 < 0 || c.L2 < 0) return fail(ctx, CW_E_INVALID, "offsets must be non-decreasing");
                    c.Lt = c.L1 + c.L2;
                    // device rows are padded to a multiple of 32 columns so every row starts 256-byte aligned (the
                    // trajectory writer's windows are aligned to absolute columns and leave in pairs)
                    c.Ltp = (c.Lt + 31) & ~int64_t(31);
                }
                layout_chunk(H, c);
                chunks.push_back(std::move(c));
                u = e1;
            }
        }
    }
    return CW_OK;
}

// stage_out -> caller arrays of a chunk whose D2H has completed
static int drain_chunk(Slot &s, Chunk &c)
{
    if (c.drained) return CW_OK;
    NvtxRange nv("cw: drain chunk (wait D2H, staged outputs -> caller arrays)");
    CW_SCK(s, cudaEventSynchronize(c.d2h_done));
    HostPool::get().run(c.drain);
    c.drain.clear();
    c.drained = true;
    return CW_OK;
}

// Everything one chunk needs, enqueued on its lane's stream: uploads (direct or staged), unpack, kernels, downloads.
static int process_chunk(HostCall &H, Slot &s, Chunk &c)
{
    NvtxRange nv("cw: enqueue chunk (stage inputs, H2D, pre-pass, sort, integrator, D2H)");
    cw_ctx *ctx = H.ctx;
    const cw_batch *B = H.B;
    const cw_result &R = H.R;
    const int64_t n = H.n, ns = H.n_src;
    cudaStream_t stream = s.lanes[c.lane];
    Arena &ar = s.lane_arena[c.lane];
    char *dev = ar.base;
    char *sin = (char *)s.stage_in[c.lane];
    char *sout = (char *)s.stage_out[c.lane];
    const int64_t m = c.m, m1 = c.m1, m2 = c.m2;
    const size_t osz = H.off32 ? 4 : 8;

    // ---- uploads
    std::vector<CopyTask> tasks;
    std::vector<std::pair<size_t, size_t>> staged;        // (offset, bytes) runs to DMA from the staging buffer
    auto up = [&](const Seg &sg, size_t sub, const void *user, size_t bytes, bool dma) -> cudaError_t {
        if (bytes == 0) return cudaSuccess;
        if (dma) return cudaMemcpyAsync(dev + sg.off + sub, user, bytes, cudaMemcpyHostToDevice, stream);
        tasks.push_back(CopyTask{sin + sg.off + sub, user, bytes});
        if (!staged.empty() && staged.back().first + pad256(staged.back().second) >= sg.off + sub &&
            staged.back().first <= sg.off + sub)
            staged.back().second = sg.off + sub + bytes - staged.back().first;     // adjacent: one DMA for both
        else
            staged.push_back({sg.off + sub, bytes});
        return cudaSuccess;
    };
    if (H.integrate)
        for (int r = 0; r < 6; r++) CW_SCK(s, up(c.u_w0, (size_t)r * m1 * 8, B->w0 + (size_t)r * ns + c.ua, m1 * 8, H.p_w0));
    CW_SCK(s, up(c.u_t1, 0, B->t1 + c.ua, m1 * 8, H.p_t1));
    if (m2) CW_SCK(s, up(c.u_idx, 0, H.idx + c.sa, m2 * 4, H.p_idx));
    if (H.classes) {
        CW_SCK(s, up(c.u_cls, 0, B->grid_class + c.ua, m1, H.p_cls));
        CW_SCK(s, up(c.u_cls, m1, B->grid_class + ns + c.sa, m2, H.p_cls));
        CW_SCK(s, up(c.u_ctab, 0, B->classes, (size_t)B->n_classes * sizeof(cw_grid_class), H.p_ctab));
    } else {
        CW_SCK(s, up(c.u_t2, 0, B->t2 + c.ua, m1 * 8, H.p_t2));
        CW_SCK(s, up(c.u_t2, m1 * 8, B->t2 + ns + c.sa, m2 * 8, H.p_t2));
        CW_SCK(s, up(c.u_dt, 0, B->dt + c.ua, m1 * 8, H.p_dt));
        CW_SCK(s, up(c.u_dt, m1 * 8, B->dt + ns + c.sa, m2 * 8, H.p_dt));
        CW_SCK(s, up(c.u_tm, 0, B->to_myr + c.ua, m1 * 8, H.p_tm));
        CW_SCK(s, up(c.u_tm, m1 * 8, B->to_myr + ns + c.sa, m2 * 8, H.p_tm));
        CW_SCK(s, up(c.u_fl, 0, B->grid_flags + c.ua, m1 * 4, H.p_fl));
        CW_SCK(s, up(c.u_fl, m1 * 4, B->grid_flags + ns + c.sa, m2 * 4, H.p_fl));
    }
    if (H.want_ev && c.K > 0) {
        const char *off = H.off32 ? (const char *)B->ev_off32 : (const char *)B->ev_off;
        CW_SCK(s, up(c.u_off1, 0, off + (size_t)c.ua * osz, (m1 + 1) * osz, H.p_off));
        if (m2) CW_SCK(s, up(c.u_off2, 0, off + (size_t)(ns + c.sa) * osz, (m2 + 1) * osz, H.p_off));
        CW_SCK(s, up(c.u_evt, 0, B->ev_time + c.e1a, c.K1 * 8, H.p_evt));
        CW_SCK(s, up(c.u_evt, c.K1 * 8, B->ev_time + c.e2a, c.K2 * 8, H.p_evt));
        if (H.integrate) {
            CW_SCK(s, up(c.u_evd, 0, B->ev_dv + 3 * c.e1a, 3 * c.K1 * 8, H.p_evd));
            CW_SCK(s, up(c.u_evd, 3 * c.K1 * 8, B->ev_dv + 3 * c.e2a, 3 * c.K2 * 8, H.p_evd));
        }
    }
    if (H.store_all) {
        CW_SCK(s, up(c.u_toff1, 0, R.traj_off + c.ua, (m1 + 1) * 8, H.p_toff));
        if (m2) CW_SCK(s, up(c.u_toff2, 0, R.traj_off + ns + c.sa, (m2 + 1) * 8, H.p_toff));
    }
    if (!tasks.empty()) {
        // the lane's staging buffer is free once the previous chunk's DMA out of it has completed
        if (s.h2d_pending[c.lane]) CW_SCK(s, cudaEventSynchronize(s.h2d_done[c.lane]));
        HostPool::get().run(tasks);
        for (const auto &run : staged)
            CW_SCK(s, cudaMemcpyAsync(dev + run.first, sin + run.first, run.second, cudaMemcpyHostToDevice, stream));
        CW_SCK(s, cudaEventRecord(s.h2d_done[c.lane], stream));
        s.h2d_pending[c.lane] = true;
    }

    // ---- kernel arguments
    Plan &pl = c.plan;
    pl.stream = stream;
    pl.sort = c.sort;
    pl.sort_bytes = c.sort_bytes;
    IntegArgs &A = pl.A;
    memset(&A, 0, sizeof A);
    A.n = m;
    A.atol = H.o->atol > 0 ? H.o->atol : 1e-10;
    A.rtol = H.o->rtol > 0 ? H.o->rtol : 1e-10;
    A.nmax = H.o->nmax;
    auto D = [&](const Seg &sg) { return dev + sg.off; };
    A.ev_tk = (double *)D(c.k_evtk);
    if (c.k_tlast.bytes) A.t_last = (double *)D(c.k_tlast);
    if (c.sort) {
        const size_t stride = pad256((size_t)m * 4);
        pl.iota = (int32_t *)D(c.k_sort4);
        pl.keys_out = (int32_t *)(D(c.k_sort4) + stride);
        pl.keys_in = (int32_t *)(D(c.k_sort4) + 2 * stride);
        pl.order = (int32_t *)(D(c.k_sort4) + 3 * stride);
        A.order = pl.order;
        pl.sort_tmp = D(c.k_sorttmp);
    }
    A.w0 = (const double *)(c.do_w0 ? D(c.k_w0) : D(c.u_w0));
    A.t1 = (const double *)(c.do_w0 ? D(c.k_t1) : D(c.u_t1));
    if (H.classes) {
        double *g = (double *)D(c.k_grid);
        A.t2 = g; A.dt = g + m; A.to_myr = g + 2 * m; A.flags = (const int32_t *)D(c.k_fl);
    } else {
        A.t2 = (const double *)D(c.u_t2); A.dt = (const double *)D(c.u_dt);
        A.to_myr = (const double *)D(c.u_tm); A.flags = (const int32_t *)D(c.u_fl);
    }
    if (H.want_ev && c.K > 0) {
        A.ev_time = (const double *)D(c.u_evt);
        A.ev_dv = H.integrate ? (const double *)D(c.u_evd) : nullptr;
        if (c.unpack) { A.ev_off = (const int64_t *)D(c.k_evoff); A.ev_base = 0; }
        else { A.ev_off = (const int64_t *)D(c.u_off1); A.ev_base = c.e1a; }
    }
    A.w_final = H.integrate ? (double *)D(c.d_wf) : nullptr;
    A.status = (int32_t *)D(c.d_st);
    A.n_nodes = (int32_t *)D(c.d_nn);
    A.ev_step = (int32_t *)D(c.d_es);
    if (H.store_all) {
        double *tp = (double *)D(c.d_tp);
        A.traj_total = c.Ltp;
        A.traj_pos = tp; A.traj_vel = tp + 3 * c.Ltp;
        A.traj_t = R.traj_t ? (double *)D(c.d_tt) : nullptr;
        if (c.unpack) { A.traj_off = (const int64_t *)D(c.k_toff); A.traj_base = 0; }
        else { A.traj_off = (const int64_t *)D(c.u_toff1); A.traj_base = c.c1a; }
    }
    unsigned long long *qs = s.meta.take<unsigned long long>(8);
    if (!qs) return slot_fail(s, CW_E_CUDA, "device arena exhausted (internal sizing error)");
    A.queue = qs;
    A.stats = qs + 1;

    if (c.unpack) {
        UnpackArgs U;
        memset(&U, 0, sizeof U);
        U.m1 = m1; U.m2 = m2; U.ua = c.ua;
        U.u_w0 = (const double *)D(c.u_w0); U.u_t1 = (const double *)D(c.u_t1);
        U.u_idx = (const int32_t *)D(c.u_idx);
        U.do_w0 = c.do_w0 ? 1 : 0;
        if (c.do_w0) { U.k_w0 = (double *)D(c.k_w0); U.k_t1 = (double *)D(c.k_t1); }
        if (H.classes) {
            U.u_cls = (const uint8_t *)D(c.u_cls);
            U.u_ctab = (const cw_grid_class *)D(c.u_ctab);
            U.n_classes = B->n_classes;
            double *g = (double *)D(c.k_grid);
            U.k_t2 = g; U.k_dt = g + m; U.k_tm = g + 2 * m; U.k_fl = (int32_t *)D(c.k_fl);
        }
        if (H.want_ev && c.K > 0) {
            U.u_off1 = D(c.u_off1); U.u_off2 = D(c.u_off2); U.off_is32 = H.off32 ? 1 : 0;
            U.k_evoff = (int64_t *)D(c.k_evoff);
        }
        if (H.store_all) {
            U.u_toff1 = (const int64_t *)D(c.u_toff1); U.u_toff2 = (const int64_t *)D(c.u_toff2);
            U.k_toff = (int64_t *)D(c.k_toff);
        }
        unpack_chunk_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, stream>>>(U);
        CW_SCK(s, cudaGetLastError());
        s.launches += 1;
    }
    int rc = run_kernels(ctx, s, H.o, pl, H.integrate, H.store_all);
    if (rc != CW_OK) return rc;

    // ---- downloads: page-locked destinations directly, pageable ones through the lane's stage_out buffer
    std::vector<std::pair<size_t, size_t>> sruns;        // (offset within the download region, bytes)
    c.drain.clear();
    auto down = [&](const Seg &sg, size_t sub, void *user, size_t bytes, bool dma) -> cudaError_t {
        if (bytes == 0 || !user) return cudaSuccess;
        if (dma) return cudaMemcpyAsync(user, dev + sg.off + sub, bytes, cudaMemcpyDeviceToHost, stream);
        const size_t rel = sg.off + sub - c.down_off;
        c.drain.push_back(CopyTask{user, sout + rel, bytes});
        if (!sruns.empty() && sruns.back().first + pad256(sruns.back().second) >= rel && sruns.back().first <= rel)
            sruns.back().second = rel + bytes - sruns.back().first;
        else
            sruns.push_back({rel, bytes});
        return cudaSuccess;
    };
    if (H.integrate) {
        for (int r = 0; r < 6; r++) {
            CW_SCK(s, down(c.d_wf, (size_t)r * m * 8, R.w_final + (size_t)r * n + c.ua, m1 * 8, H.p_wf));
            CW_SCK(s, down(c.d_wf, ((size_t)r * m + m1) * 8, R.w_final + (size_t)r * n + ns + c.sa, m2 * 8, H.p_wf));
        }
        CW_SCK(s, down(c.d_st, 0, R.status + c.ua, m1 * 4, H.p_st));
        CW_SCK(s, down(c.d_st, m1 * 4, R.status + ns + c.sa, m2 * 4, H.p_st));
    }
    if (R.n_nodes) {
        CW_SCK(s, down(c.d_nn, 0, R.n_nodes + c.ua, m1 * 4, H.p_nn));
        CW_SCK(s, down(c.d_nn, m1 * 4, R.n_nodes + ns + c.sa, m2 * 4, H.p_nn));
    }
    if (R.ev_step && H.want_ev && c.K > 0) {
        CW_SCK(s, down(c.d_es, 0, R.ev_step + c.e1a, c.K1 * 4, H.p_es));
        CW_SCK(s, down(c.d_es, c.K1 * 4, R.ev_step + c.e2a, c.K2 * 4, H.p_es));
    }
    if (H.store_all && c.Lt > 0) {
        const size_t tc = (size_t)H.total_cols;
        for (int r = 0; r < 3; r++) {
            CW_SCK(s, down(c.d_tp, (size_t)r * c.Ltp * 8, R.traj_pos + r * tc + c.c1a, c.L1 * 8, H.p_tp));
            CW_SCK(s, down(c.d_tp, ((size_t)r * c.Ltp + c.L1) * 8, R.traj_pos + r * tc + c.c2a, c.L2 * 8, H.p_tp));
            CW_SCK(s, down(c.d_tp, (size_t)(3 + r) * c.Ltp * 8, R.traj_vel + r * tc + c.c1a, c.L1 * 8, H.p_tv));
            CW_SCK(s, down(c.d_tp, ((size_t)(3 + r) * c.Ltp + c.L1) * 8, R.traj_vel + r * tc + c.c2a, c.L2 * 8, H.p_tv));
        }
        if (R.traj_t) {
            CW_SCK(s, down(c.d_tt, 0, R.traj_t + c.c1a, c.L1 * 8, H.p_tt));
            CW_SCK(s, down(c.d_tt, c.L1 * 8, R.traj_t + c.c2a, c.L2 * 8, H.p_tt));
        }
    }
    if (!c.drain.empty()) {
        for (const auto &run : sruns)
            CW_SCK(s, cudaMemcpyAsync(sout + run.first, dev + c.down_off + run.first, run.second, cudaMemcpyDeviceToHost, stream));
        CW_SCK(s, cudaEventRecord(c.d2h_done, stream));
        c.drained = false;
    }
    return CW_OK;
}

// all chunks of one device slot, in order, on the calling thread
static int run_slot_host(HostCall &H, int d, std::vector<Chunk *> &mine)
{
    cw_ctx *ctx = H.ctx;
    Slot &s = ctx->slots[d];
    s.error.clear();
    s.launches = 0;
    if (mine.empty()) return CW_OK;
    NvtxRange nv("cw: device slot (host thread)");
    const double tr_start = H.trace ? H.ms_since_call() : 0.0;
    CW_SCK(s, cudaSetDevice(s.dev));
    // ---- size and reserve the lane arenas and staging buffers once, for all chunks of the call
    size_t need[Slot::N_LANES] = {0, 0, 0}, need_in[Slot::N_LANES] = {0, 0, 0}, need_out[Slot::N_LANES] = {0, 0, 0};
    const bool any_in_staged = !(H.p_w0 && H.p_t1 && H.p_t2 && H.p_dt && H.p_tm && H.p_fl && H.p_cls && H.p_ctab &&
                                 H.p_idx && H.p_off && H.p_evt && H.p_evd && H.p_toff);
    const bool any_out_staged = !(H.p_wf && H.p_st && H.p_nn && H.p_es && H.p_tp && H.p_tv && H.p_tt);
    for (Chunk *c : mine) {
        need[c->lane] = std::max(need[c->lane], c->total_bytes);
        if (any_in_staged) need_in[c->lane] = std::max(need_in[c->lane], c->up_bytes);
        if (any_out_staged) need_out[c->lane] = std::max(need_out[c->lane], c->down_bytes);
    }
    for (int l = 0; l < Slot::N_LANES; l++) {
        if (need[l] == 0) continue;
        if (need[l] + 4096 > s.lane_arena[l].cap) cudaStreamSynchronize(s.lanes[l]);
        CW_SCK(s, s.lane_arena[l].reserve(need[l] + 4096));
        if (need_in[l] > s.stage_in_cap[l]) {
            pinned_free(s.stage_in[l]);
            s.stage_in[l] = pinned_alloc(need_in[l] + need_in[l] / 8);
            s.stage_in_cap[l] = s.stage_in[l] ? need_in[l] + need_in[l] / 8 : 0;
            if (!s.stage_in[l]) return slot_fail(s, CW_E_CUDA, "cannot allocate the page-locked staging buffer (inputs)");
        }
        if (need_out[l] > s.stage_out_cap[l]) {
            pinned_free(s.stage_out[l]);
            s.stage_out[l] = pinned_alloc(need_out[l] + need_out[l] / 8);
            s.stage_out_cap[l] = s.stage_out[l] ? need_out[l] + need_out[l] / 8 : 0;
            if (!s.stage_out[l]) return slot_fail(s, CW_E_CUDA, "cannot allocate the page-locked staging buffer (outputs)");
        }
        s.h2d_pending[l] = false;
    }
    const size_t meta_need = Arena::padded(8, 8) * mine.size() + 4096;
    if (meta_need > s.meta.cap)
        for (int l = 0; l < Slot::N_LANES; l++) cudaStreamSynchronize(s.lanes[l]);
    CW_SCK(s, s.meta.reserve(meta_need));
    CW_SCK(s, s.ensure_events(4 * mine.size()));
    const double tr_reserved = H.trace ? H.ms_since_call() : 0.0;
    // ---- enqueue chunk after chunk; a chunk re-uses the buffers of the one three places back, which must have been
    // drained by then
    int rc = CW_OK;
    for (size_t k = 0; k < mine.size() && rc == CW_OK; k++) {
        Chunk &c = *mine[k];
        c.plan.ev_t0 = s.events[4 * k]; c.plan.ev_k0 = s.events[4 * k + 1]; c.plan.ev_k1 = s.events[4 * k + 2];
        c.d2h_done = s.events[4 * k + 3];
        if (k >= (size_t)Slot::N_LANES) rc = drain_chunk(s, *mine[k - Slot::N_LANES]);
        // ... and whatever else has already arrived (keeps the tail short)
        for (size_t j = 0; j < k && rc == CW_OK; j++)
            if (!mine[j]->drained && cudaEventQuery(mine[j]->d2h_done) == cudaSuccess) rc = drain_chunk(s, *mine[j]);
        if (rc == CW_OK) rc = process_chunk(H, s, c);
    }
    const double tr_enqueued = H.trace ? H.ms_since_call() : 0.0;
    for (size_t k = 0; k < mine.size() && rc == CW_OK; k++) rc = drain_chunk(s, *mine[k]);
    const double tr_drained = H.trace ? H.ms_since_call() : 0.0;
    // ---- wait for the lanes (also on failure: nothing may still be reading or writing caller memory on return)
    for (int l = 0; l < Slot::N_LANES; l++) {
        const cudaError_t e = cudaStreamSynchronize(s.lanes[l]);
        if (e != cudaSuccess && rc == CW_OK) {
            s.error = std::string("cudaStreamSynchronize failed: ") + cudaGetErrorString(e);
            rc = CW_E_CUDA;
        }
    }
    cudaGetLastError();
    if (H.trace)
        fprintf(stderr, "[cw] slot %d (device %d): %zu chunks; host thread started %.3f, buffers reserved %.3f, all chunks "
                        "enqueued %.3f, drained %.3f, lanes idle %.3f ms after the call\n", d, s.dev, mine.size(), tr_start,
                tr_reserved, tr_enqueued, tr_drained, H.ms_since_call());
    return rc;
}

static int integrate_host(cw_ctx *ctx, const cw_opts *o, const cw_batch *B, const cw_result *R, bool integrate,
                          int32_t *n_nodes_out, int32_t *ev_step_out)
{
    NvtxRange nv(integrate ? "cw_integrate (CW_MEM_HOST)" : "cw_count_nodes (CW_MEM_HOST)");
    HostCall H;
    H.ctx = ctx; H.o = o; H.B = B;
    H.trace = getenv("CW_DEBUG") != nullptr;
    H.t_call = std::chrono::steady_clock::now();
    memset(&H.R, 0, sizeof H.R);
    if (R) H.R = *R;
    if (!integrate) { H.R.n_nodes = n_nodes_out; H.R.ev_step = ev_step_out; }
    H.integrate = integrate;
    H.store_all = integrate && R->traj_off;
    H.n = B->n_orbits;
    H.n_src = (B->n_w0 > 0) ? B->n_w0 : H.n;
    H.n_sec = H.n - H.n_src;
    H.idx = H.n_sec > 0 ? B->w0_index : nullptr;
    if (!integrate && H.n_sec > 0) return fail(ctx, CW_E_UNSUPPORTED, "cw_count_nodes does not take w0_index (pass the plain t1 column)");
    H.off32 = B->ev_off32 != nullptr;
    H.has_ev = (B->ev_off || B->ev_off32) && H.ev_at(H.n) > 0;
    H.want_ev = H.has_ev && (integrate || ev_step_out);
    if (H.has_ev && H.want_ev && (!B->ev_time || (integrate && !B->ev_dv)))
        return fail(ctx, CW_E_INVALID, "ev_off lists events but ev_time / ev_dv is null");
    H.classes = B->grid_class != nullptr;
    H.total_cols = H.store_all ? R->traj_off[H.n] : 0;
    if (H.n_sec > 0) {
        // cheap end checks; an unsorted index is caught per orbit by the unpack kernel (status CW_ORBIT_INCONSISTENT)
        if (H.n_src < 1 || H.idx[0] < 0 || H.idx[H.n_sec - 1] >= H.n_src || H.idx[0] > H.idx[H.n_sec - 1])
            return fail(ctx, CW_E_INVALID, "w0_index must be non-decreasing with entries in [0, n_w0)");
    }
    auto pin = [](const void *p) { return pointer_is_dma_able(p); };
    const bool force_stage = getenv("CW_FORCE_STAGING") != nullptr;     // test hook: treat every array as pageable
    auto P = [&](const void *p) { return !force_stage && pin(p); };
    H.p_w0 = P(B->w0); H.p_t1 = P(B->t1); H.p_t2 = P(B->t2); H.p_dt = P(B->dt); H.p_tm = P(B->to_myr);
    H.p_fl = P(B->grid_flags); H.p_cls = P(B->grid_class); H.p_ctab = P(B->classes); H.p_idx = P(H.idx);
    H.p_off = P(H.off32 ? (const void *)B->ev_off32 : (const void *)B->ev_off);
    H.p_evt = P(B->ev_time); H.p_evd = P(B->ev_dv); H.p_toff = P(H.R.traj_off);
    H.p_wf = P(H.R.w_final); H.p_st = P(H.R.status); H.p_nn = P(H.R.n_nodes); H.p_es = P(H.R.ev_step);
    H.p_tp = P(H.R.traj_pos); H.p_tv = P(H.R.traj_vel); H.p_tt = P(H.R.traj_t);

    DeviceGuard guard;
    std::vector<Chunk> chunks;
    int rc = make_chunks(H, chunks);
    if (rc != CW_OK) return rc;
    const int nslots = (int)ctx->slots.size();
    std::vector<std::vector<Chunk *>> per_slot(nslots);
    for (Chunk &c : chunks) per_slot[c.slot].push_back(&c);
    // a device-mode call left running (no_sync) must not see its scratch move
    if (ctx->slots[0].has_pending) {
        cudaSetDevice(ctx->slots[0].dev);
        cudaEventSynchronize(ctx->slots[0].pending);
        ctx->slots[0].has_pending = false;
    }
    // ---- one host thread per device slot; the caller's thread takes the first busy slot
    std::vector<int> rcs(nslots, CW_OK);
    std::vector<int> busy;
    int first = -1;
    if ((int)ctx->workers.size() < nslots) ctx->workers.resize(nslots);
    for (int d = 0; d < nslots; d++) {
        if (per_slot[d].empty()) continue;
        if (first < 0) { first = d; continue; }
        if (!ctx->workers[d]) ctx->workers[d].reset(new SlotWorker());
        ctx->workers[d]->submit([&, d] { rcs[d] = run_slot_host(H, d, per_slot[d]); });
        busy.push_back(d);
    }
    if (first >= 0) rcs[first] = run_slot_host(H, first, per_slot[first]);
    for (int d : busy) ctx->workers[d]->wait();
    if (H.trace) fprintf(stderr, "[cw] %zu chunks planned and every slot joined %.3f ms after the call\n", chunks.size(), H.ms_since_call());
    ctx->last_launches = 0;
    for (int d = 0; d < nslots; d++) {
        ctx->last_launches += ctx->slots[d].launches;
        if (rcs[d] != CW_OK && rc == CW_OK) { rc = rcs[d]; ctx->last_error = ctx->slots[d].error; }
    }
    if (rc != CW_OK) return rc;
    // ---- statistics and kernel times
    int64_t stats[5] = {0, 0, 0, 0, 0};
    const bool want_stats = integrate;
    ctx->last_max_attempts = 0;
    for (int d = 0; d < nslots; d++) {
        if (per_slot[d].empty()) continue;
        Slot &s = ctx->slots[d];
        cudaSetDevice(s.dev);
        const size_t mstride = Arena::padded(8, 8) / 8;       // one padded block of 8 counters per chunk
        std::vector<unsigned long long> hs(mstride * per_slot[d].size(), 0ull);
        if (want_stats)
            CW_CK(ctx, cudaMemcpy(hs.data(), s.meta.base, hs.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < per_slot[d].size(); k++) {
            const Plan &pl = per_slot[d][k]->plan;
            for (int q = 0; q < 5; q++) stats[q] += (int64_t)hs[mstride * k + 1 + q];
            ctx->last_max_attempts = std::max(ctx->last_max_attempts, (int64_t)hs[mstride * k + 6]);
            if (pl.timed) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, pl.ev_k0, pl.ev_k1) == cudaSuccess) s.last_kernel_ms += ms;
                if (cudaEventElapsedTime(&ms, pl.ev_t0, pl.ev_k1) == cudaSuccess) s.last_total_ms += ms;
            }
        }
    }
    if (getenv("CW_DEBUG") && first >= 0 && per_slot[first][0]->plan.timed) {
        // chunk timeline relative to the first chunk's first kernel: pre-pass start, integrator start / end
        const Plan &p0 = per_slot[first][0]->plan;
        cudaSetDevice(ctx->slots[first].dev);
        for (size_t k = 0; k < per_slot[first].size(); k++) {
            const Chunk &c = *per_slot[first][k];
            if (!c.plan.timed) continue;
            float a = 0.f, b = 0.f, e = 0.f;
            cudaEventElapsedTime(&a, p0.ev_t0, c.plan.ev_t0);
            cudaEventElapsedTime(&b, p0.ev_t0, c.plan.ev_k0);
            cudaEventElapsedTime(&e, p0.ev_t0, c.plan.ev_k1);
            fprintf(stderr, "[cw] chunk %2zu lane %d orbits %8lld (+%lld): prepass %8.3f  kernel %8.3f .. %8.3f ms\n", k,
                    c.lane, (long long)c.m1, (long long)c.m2, a, b, e);
        }
    }
    if (integrate && R->stats) memcpy(R->stats, stats, 4 * sizeof(int64_t));
    if (integrate && R->n_failed) *R->n_failed = stats[4];
    return CW_OK;
}

} // namespace

static int integrate_impl(cw_ctx *ctx, const cw_opts *o, const cw_batch *B, const cw_result *R, bool integrate,
                          int32_t *n_nodes_out, int32_t *ev_step_out)
{
    int rc = validate(ctx, o, B, integrate);
    if (rc != CW_OK) return rc;
    if (integrate && !ctx->has_pot) return fail(ctx, CW_E_NO_POTENTIAL, "cw_set_potential has not been called");
    if (integrate && (!R || !R->w_final || !R->status)) return fail(ctx, CW_E_INVALID, "w_final and status are required");
    ctx->last_launches = 0;
    for (Slot &s : ctx->slots) { s.last_kernel_ms = s.last_total_ms = 0.0; }
    if (B->n_orbits == 0) return CW_OK;
    if (integrate && R->traj_off && (!R->traj_pos || !R->traj_vel)) return fail(ctx, CW_E_INVALID, "traj_pos / traj_vel missing");
    if (o->mem_space == CW_MEM_DEVICE) return integrate_device(ctx, o, B, R, integrate, n_nodes_out, ev_step_out);
    return integrate_host(ctx, o, B, R, integrate, n_nodes_out, ev_step_out);
}

extern "C" int cw_count_nodes(cw_ctx *ctx, const cw_opts *opts, const cw_batch *batch, int32_t *n_nodes, int32_t *ev_step)
{
    if (!n_nodes) return fail(ctx, CW_E_INVALID, "n_nodes is required");
    return integrate_impl(ctx, opts, batch, nullptr, false, n_nodes, ev_step);
}

extern "C" int cw_integrate(cw_ctx *ctx, const cw_opts *opts, const cw_batch *batch, const cw_result *result)
{
    return integrate_impl(ctx, opts, batch, result, true, nullptr, nullptr);
}

extern "C" int64_t cw_last_launch_count(const cw_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

extern "C" int64_t cw_last_max_attempts(const cw_ctx *ctx) { return ctx ? ctx->last_max_attempts : 0; }

extern "C" double cw_last_kernel_ms(const cw_ctx *ctx, int dev)
{
    if (!ctx || dev < 0 || dev >= (int)ctx->slots.size()) return 0.0;
    return ctx->slots[dev].last_kernel_ms;
}

extern "C" double cw_last_total_kernel_ms(const cw_ctx *ctx, int dev)
{
    if (!ctx || dev < 0 || dev >= (int)ctx->slots.size()) return 0.0;
    return ctx->slots[dev].last_total_ms;
}

// ---------------------------------------------------------------------------------------------
// pointwise potential entry points

static int pointwise(cw_ctx *ctx, int mode, int mem_space, int64_t n, const double *q, double *out, const double *tt = nullptr)
{
    if (!ctx || !q || !out || n < 0) return fail(ctx, CW_E_INVALID, "bad arguments");
    if (!ctx->has_pot) return fail(ctx, CW_E_NO_POTENTIAL, "cw_set_potential has not been called");
    if (ctx->slots.empty()) return CW_E_NO_DEVICE;
    if (n == 0) return CW_OK;
    Slot &s = ctx->slots[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s.dev);
    LaunchCfg cfg{s.sm_count, s.stream};
    const int64_t n_out = (mode == 0) ? 3 * n : n;
    ctx->last_launches = 1;
    if (mem_space == CW_MEM_DEVICE) {
        CW_CK(ctx, launch_pointwise(ctx->pot, mode, n, q, tt, out, cfg));
    } else {
        CW_CK(ctx, s.arena.reserve(Arena::padded(3 * n, 8) + Arena::padded(n_out, 8) + Arena::padded(n, 8) + 4096));
        double *dq = s.arena.take<double>(3 * n), *dout = s.arena.take<double>(n_out);
        double *dt = tt ? s.arena.take<double>(n) : nullptr;
        CW_CK(ctx, cudaMemcpyAsync(dq, q, 3 * n * 8, cudaMemcpyHostToDevice, s.stream));
        if (tt) CW_CK(ctx, cudaMemcpyAsync(dt, tt, n * 8, cudaMemcpyHostToDevice, s.stream));
        CW_CK(ctx, launch_pointwise(ctx->pot, mode, n, dq, dt, dout, cfg));
        CW_CK(ctx, cudaMemcpyAsync(out, dout, n_out * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    CW_CK(ctx, cudaStreamSynchronize(s.stream));
    cudaSetDevice(prev);
    return CW_OK;
}

extern "C" int cw_gradient(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *grad)
{
    return pointwise(ctx, 0, mem_space, n, q, grad);
}
extern "C" int cw_potential_value(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *phi)
{
    return pointwise(ctx, 1, mem_space, n, q, phi);
}
extern "C" int cw_circular_velocity(cw_ctx *ctx, int mem_space, int64_t n, const double *q, double *vcirc)
{
    return pointwise(ctx, 2, mem_space, n, q, vcirc);
}

extern "C" int cw_pointwise_t(cw_ctx *ctx, int what, int mem_space, int64_t n, const double *q, const double *t_myr, double *out)
{
    if (what < 0 || what > 2) return fail(ctx, CW_E_INVALID, "cw_pointwise_t: what must be CW_GRADIENT, CW_VALUE or CW_VCIRC");
    return pointwise(ctx, what, mem_space, n, q, out, t_myr);
}

// ---------------------------------------------------------------------------------------------
// diagnostics

extern "C" double cw_measure_fp64_peak(cw_ctx *ctx, int dev, double ms_target)
{
    if (!ctx || dev < 0 || dev >= (int)ctx->slots.size()) return -1.0;
    Slot &s = ctx->slots[dev];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s.dev);
    LaunchCfg cfg{s.sm_count, s.stream};
    double result = -1.0;
    do {
        if (s.arena.reserve(4096) != cudaSuccess) break;
        double *sink = s.arena.take<double>(8);
        int blocks = 0, threads = 0, iters = 20000;
        float ms = 0.f;
        // warm-up + calibration pass, then the timed pass sized for ~ms_target
        for (int pass = 0; pass < 2; pass++) {
            if (cudaEventRecord(s.ev_k0, s.stream) != cudaSuccess) break;
            if (launch_dfma_peak(sink, iters, cfg, &blocks, &threads) != cudaSuccess) break;
            if (cudaEventRecord(s.ev_k1, s.stream) != cudaSuccess) break;
            if (cudaStreamSynchronize(s.stream) != cudaSuccess) break;
            if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) != cudaSuccess) break;
            if (pass == 0 && ms > 0.f) {
                double scale = (ms_target > 0 ? ms_target : 50.0) / ms;
                if (scale > 200.0) scale = 200.0;
                if (scale < 0.05) scale = 0.05;
                iters = (int)(iters * scale);
                if (iters < 1000) iters = 1000;
            }
        }
        if (ms > 0.f) result = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads / (ms * 1e-3) / 1e12;
    } while (0);
    cudaGetLastError();
    cudaSetDevice(prev);
    return result;
}

extern "C" double cw_measure_scatter_write(cw_ctx *ctx, int dev, int64_t streams, int64_t stream_bytes, int run_bytes,
                                           int resident_threads_per_sm)
{
    int rb = run_bytes < 0 ? -run_bytes : run_bytes;      // negative: with L2 eviction-priority hints
    if (run_bytes == 1128) rb = 128;                      // one thread per stream, 128-byte runs as four 32-byte stores
    if (!ctx || dev < 0 || dev >= (int)ctx->slots.size() || rb < 16 || rb > 512 || (rb & 15) || stream_bytes % rb)
        return -1.0;
    Slot &s = ctx->slots[dev];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s.dev);
    LaunchCfg cfg{s.sm_count, s.stream};
    double result = -1.0;
    do {
        const size_t bytes = (size_t)streams * (size_t)stream_bytes;
        if (s.arena.reserve(bytes + 4096) != cudaSuccess) break;
        double2 *base = s.arena.take<double2>(bytes / 16);
        if (!base) break;
        int blocks = s.sm_count * std::max(1, resident_threads_per_sm / 256);
        float ms = 0.f;
        bool ok = true;
        for (int pass = 0; pass < 2 && ok; pass++) {
            ok = ok && cudaEventRecord(s.ev_k0, s.stream) == cudaSuccess;
            ok = ok && launch_scatter_write(base, streams, stream_bytes, run_bytes, blocks, cfg) == cudaSuccess;
            ok = ok && cudaEventRecord(s.ev_k1, s.stream) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(s.stream) == cudaSuccess;
            ok = ok && cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess;
        }
        if (ok && ms > 0.f) result = (double)bytes / (ms * 1e-3) / 1e9;
    } while (0);
    cudaGetLastError();
    cudaSetDevice(prev);
    return result;
}

extern "C" int cw_selftest_math(cw_ctx *ctx, int64_t n, double *out)
{
    if (!ctx || !out || ctx->slots.empty()) return CW_E_INVALID;
    Slot &s = ctx->slots[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s.dev);
    LaunchCfg cfg{s.sm_count, s.stream};
    CW_CK(ctx, s.arena.reserve(4096));
    double *d = s.arena.take<double>(8);
    CW_CK(ctx, launch_math_selftest(n, d, cfg));
    CW_CK(ctx, cudaMemcpyAsync(out, d, 7 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CW_CK(ctx, cudaStreamSynchronize(s.stream));
    cudaSetDevice(prev);
    return CW_OK;
}
