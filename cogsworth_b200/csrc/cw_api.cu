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
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "cw_kernels.cuh"

using namespace cw;

namespace {

// restores the caller's current device on every exit path (the CW_CK macro returns early on CUDA errors)
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
    std::vector<cudaEvent_t> events;        // per-shard timing events (grow-only pool)
    Arena arena;                            // diagnostics / pointwise entry points
    // Integration scratch, ONE ARENA PER STREAM LANE: the shards of a lane run one after another on the lane's
    // stream, so each of them re-uses the lane's buffers (stream order is the only synchronisation needed).
    // Device memory of a host-buffer call is therefore bounded by N_LANES chunks, not by the size of the batch:
    // trajectories larger than HBM stream through.
    Arena lane_arena[N_LANES];
    Arena meta;                             // queue counter + statistics of every shard of a call (64 B each): NOT
                                            // re-used within a call, so they can be read back after the last shard
    double last_kernel_ms = 0.0, last_total_ms = 0.0;
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

struct cw_ctx {
    std::vector<Slot> slots;
    PotParams pot;
    bool has_pot = false;
    std::string last_error;
    int64_t last_launches = 0;
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
    if (P.kind == POT_GENERAL)
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

namespace {

// One unit of work: a contiguous range of orbits on one device slot, on one of its streams.
struct Shard {
    int slot = 0, lane = 0; // device slot, stream lane within the slot
    int64_t a = 0, b = 0;   // orbit range
    int64_t ea = 0, eb = 0; // event range
    int64_t ta = 0, tb = 0; // trajectory column range
};

// size of the CUB temp storage for sorting m (key, value) pairs
size_t sort_temp_bytes(int64_t m)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                              (const int32_t *)nullptr, (int32_t *)nullptr, (int)m);
    return bytes;
}

constexpr int64_t SORT_THRESHOLD = 4096;
// Host-buffer calls are cut into chunks that rotate over N_LANES streams, so the H2D copy of chunk c+1
// and the D2H copy of chunk c-1 overlap the kernels of chunk c: about 8 chunks per device, between
// 2^16 orbits (below that launch overheads show) and 2^20 (bounds the exposed first copy).
constexpr int64_t HOST_CHUNK_MAX = 1 << 20, HOST_CHUNK_MIN = 1 << 16;
constexpr int64_t TRAJ_SHARD_COLS = (int64_t)1 << 26;   // trajectory columns per host-mode shard (store_all)
static int64_t traj_shard_cols()
{
    const char *e = getenv("CW_TRAJ_SHARD_COLS");       // test hook: force many small shards
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (int64_t)v : TRAJ_SHARD_COLS;
}
// Q > 0: the number of orbits one full grid of the kernel holds (persistent lanes).  Chunks are whole multiples of it,
// so no chunk ends with a wave that fills a fraction of the GPU (10^6 orbits in LM10 at Q = 113 664: 127k-orbit
// chunks = 1.12 waves each cost a third of the end-to-end throughput).
static int64_t host_chunk(int64_t n_on_device, int64_t Q)
{
    int64_t c = (n_on_device + 7) / 8;
    if (Q <= 0 || getenv("CW_NO_WAVE_CHUNKS")) {
        c = (c + 4095) & ~int64_t(4095);
        return std::min(HOST_CHUNK_MAX, std::max(HOST_CHUNK_MIN, c));
    }
    c = ((c + Q - 1) / Q) * Q;
    const int64_t cmax = std::max<int64_t>(Q, (HOST_CHUNK_MAX / Q) * Q);
    return std::min(cmax, std::max(Q, c));
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

} // namespace

// bytes of device scratch one shard needs (host mode: everything; device mode: scratch only)
static size_t shard_bytes(const cw_opts *o, const cw_result *R, const Shard &sh, bool integrate, Plan &pl)
{
    const bool dev_mode = o->mem_space == CW_MEM_DEVICE;
    const int64_t m = sh.b - sh.a, K = sh.eb - sh.ea, Lt = sh.tb - sh.ta;
    const bool store_all = integrate && R && R->traj_off;
    pl.sort = integrate && m >= SORT_THRESHOLD;
    pl.sort_bytes = pl.sort ? sort_temp_bytes(m) : 0;
    size_t need = 0;
    auto add = [&](size_t count, size_t elt) { need += Arena::padded(count, elt); };
    add(K + 1, sizeof(double));          // ev_tk
    if (integrate && o->integrator == CW_DOP853_DENSE) add(m, sizeof(double));   // t_last
    add(1, 64);                          // queue + stats
    if (pl.sort) { add(m, 4); add(m, 4); add(m, 4); add(m, 4); add(pl.sort_bytes, 1); }
    if (!dev_mode) {
        add(6 * m, 8); add(4 * m, 8); add(m, 4); add(m + 1, 8); add(K + 1, 8); add(3 * K + 3, 8);
        add(m, 4); add(K + 1, 4);               // n_nodes, ev_step
        add(6 * m, 8); add(m, 4);               // w_final, status
        if (store_all) { add(m + 1, 8); add(6 * (Lt + 16), 8); add(Lt + 1, 8); }
    } else {
        add(m, 4); add(K + 1, 4); add(m, 4);    // n_nodes / ev_step / status if the caller passed NULL
    }
    return need + 1024;
}

// Carve the shard's buffers out of the slot arena, fill its IntegArgs and enqueue the H2D copies.
static int plan_shard(cw_ctx *ctx, Slot &s, const cw_opts *o, const cw_batch *B, const cw_result *R,
                      const Shard &sh, bool integrate, Plan &pl)
{
    const bool dev_mode = o->mem_space == CW_MEM_DEVICE;
    const int64_t m = sh.b - sh.a, K = sh.eb - sh.ea, Lt = sh.tb - sh.ta;
    const bool store_all = integrate && R && R->traj_off;
    cudaStream_t stream = pl.stream;
    Arena &ar = s.lane_arena[sh.lane];
    ar.used = 0;                            // the previous shard of this lane precedes us in stream order

    IntegArgs &A = pl.A;
    memset(&A, 0, sizeof A);
    A.n = m;
    A.atol = o->atol > 0 ? o->atol : 1e-10;
    A.rtol = o->rtol > 0 ? o->rtol : 1e-10;
    A.nmax = o->nmax;
    A.ev_tk = ar.take<double>(K + 1);
    if (integrate && o->integrator == CW_DOP853_DENSE) {
        A.t_last = ar.take<double>(m);
        if (!A.t_last) return fail(ctx, CW_E_CUDA, "device arena exhausted (internal sizing error)");
    }
    unsigned long long *qs = s.meta.take<unsigned long long>(8);
    A.queue = qs;
    A.stats = qs + 1;
    if (pl.sort) {
        pl.iota = ar.take<int32_t>(m);
        pl.keys_out = ar.take<int32_t>(m);
        pl.keys_in = ar.take<int32_t>(m);
        pl.order = ar.take<int32_t>(m);
        A.order = pl.order;
        pl.sort_tmp = ar.take<char>(pl.sort_bytes);
    }
    if (!dev_mode) {
        double *w0 = ar.take<double>(6 * m);
        double *tt = ar.take<double>(4 * m);
        int32_t *fl = ar.take<int32_t>(m);
        int64_t *eo = ar.take<int64_t>(m + 1);
        double *et = ar.take<double>(K + 1);
        double *ed = ar.take<double>(3 * K + 3);
        A.n_nodes = ar.take<int32_t>(m);
        A.ev_step = ar.take<int32_t>(K + 1);
        A.w_final = ar.take<double>(6 * m);
        A.status = ar.take<int32_t>(m);
        if (!w0 || !tt || !fl || !eo || !et || !ed || !A.w_final || !A.status)
            return fail(ctx, CW_E_CUDA, "device arena exhausted (internal sizing error)");
        A.w0 = w0; A.t1 = tt; A.t2 = tt + m; A.dt = tt + 2 * m; A.to_myr = tt + 3 * m; A.flags = fl;
        for (int c = 0; c < 6; c++)
            CW_CK(ctx, cudaMemcpyAsync(w0 + c * m, B->w0 + c * B->n_orbits + sh.a, m * 8, cudaMemcpyHostToDevice, stream));
        CW_CK(ctx, cudaMemcpyAsync(tt, B->t1 + sh.a, m * 8, cudaMemcpyHostToDevice, stream));
        CW_CK(ctx, cudaMemcpyAsync(tt + m, B->t2 + sh.a, m * 8, cudaMemcpyHostToDevice, stream));
        CW_CK(ctx, cudaMemcpyAsync(tt + 2 * m, B->dt + sh.a, m * 8, cudaMemcpyHostToDevice, stream));
        CW_CK(ctx, cudaMemcpyAsync(tt + 3 * m, B->to_myr + sh.a, m * 8, cudaMemcpyHostToDevice, stream));
        CW_CK(ctx, cudaMemcpyAsync(fl, B->grid_flags + sh.a, m * 4, cudaMemcpyHostToDevice, stream));
        if (B->ev_off && K > 0) {
            CW_CK(ctx, cudaMemcpyAsync(eo, B->ev_off + sh.a, (m + 1) * 8, cudaMemcpyHostToDevice, stream));
            CW_CK(ctx, cudaMemcpyAsync(et, B->ev_time + sh.ea, K * 8, cudaMemcpyHostToDevice, stream));
            CW_CK(ctx, cudaMemcpyAsync(ed, B->ev_dv + 3 * sh.ea, 3 * K * 8, cudaMemcpyHostToDevice, stream));
            A.ev_off = eo; A.ev_base = sh.ea; A.ev_time = et; A.ev_dv = ed;
        }
        if (store_all) {
            // device rows are padded to a multiple of 16 columns so every row starts 128-byte aligned
            // (the trajectory writer's bulk stores need 16-byte aligned destinations)
            const int64_t Ltp = (Lt + 15) & ~int64_t(15);
            int64_t *to = ar.take<int64_t>(m + 1);
            double *tp = ar.take<double>(6 * Ltp);
            double *ttm = ar.take<double>(Lt + 1);
            if (!to || !tp || !ttm) return fail(ctx, CW_E_CUDA, "device arena exhausted (internal sizing error)");
            CW_CK(ctx, cudaMemcpyAsync(to, R->traj_off + sh.a, (m + 1) * 8, cudaMemcpyHostToDevice, stream));
            A.traj_off = to; A.traj_base = sh.ta; A.traj_total = Ltp;
            A.traj_pos = tp; A.traj_vel = tp + 3 * Ltp; A.traj_t = R->traj_t ? ttm : nullptr;
        }
    } else {
        A.w0 = B->w0; A.t1 = B->t1; A.t2 = B->t2; A.dt = B->dt; A.to_myr = B->to_myr; A.flags = B->grid_flags;
        if (B->ev_off && K > 0) { A.ev_off = B->ev_off; A.ev_base = 0; A.ev_time = B->ev_time; A.ev_dv = B->ev_dv; }
        int32_t *nn = ar.take<int32_t>(m), *es = ar.take<int32_t>(K + 1), *st = ar.take<int32_t>(m);
        A.n_nodes = (R && R->n_nodes) ? R->n_nodes : nn;
        A.ev_step = (R && R->ev_step) ? R->ev_step : es;
        A.status = (R && R->status) ? R->status : st;
        A.w_final = R ? R->w_final : nullptr;
        if (store_all) {
            A.traj_off = R->traj_off; A.traj_base = 0; A.traj_total = Lt;
            A.traj_pos = R->traj_pos; A.traj_vel = R->traj_vel; A.traj_t = R->traj_t;
        }
    }
    if (!A.ev_tk || !A.queue || !A.n_nodes || !A.ev_step || !A.status)
        return fail(ctx, CW_E_CUDA, "device arena exhausted (internal sizing error)");
    return CW_OK;
}

static int run_shard(cw_ctx *ctx, Slot &s, const cw_opts *o, Plan &pl, bool integrate, bool store_all)
{
    cudaStream_t stream = pl.stream;
    LaunchCfg cfg{s.sm_count, stream};
    IntegArgs &A = pl.A;
    CW_CK(ctx, cudaMemsetAsync(A.queue, 0, 64, stream));
    CW_CK(ctx, cudaEventRecord(pl.ev_t0, stream));
    CW_CK(ctx, launch_grid_prepass(A, pl.sort ? pl.iota : nullptr, pl.sort ? pl.keys_in : nullptr, cfg));
    ctx->last_launches += 1;
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
                CW_CK(ctx, launch_dop853_pilot(ctx->pot, A, cfg));
                ctx->last_launches += 1;
                CW_CK(ctx, cudaMemsetAsync(A.queue, 0, 8, stream));
                A.order = pl.order;
                keys = pl.keys_in;
            }
            CW_CK(ctx, cub::DeviceRadixSort::SortPairsDescending(pl.sort_tmp, pl.sort_bytes, keys,
                                                                 pl.keys_out, (const int32_t *)pl.iota, A.order,
                                                                 (int)A.n, 0, 32, stream));
        }
        CW_CK(ctx, cudaEventRecord(pl.ev_k0, stream));
        if (o->integrator == CW_LEAPFROG) CW_CK(ctx, launch_leapfrog(ctx->pot, A, store_all, cfg));
        else CW_CK(ctx, launch_dop853(ctx->pot, A, store_all, o->integrator == CW_DOP853_DENSE, cfg));
        ctx->last_launches += 1;
        CW_CK(ctx, cudaEventRecord(pl.ev_k1, stream));
        pl.timed = true;
        if (store_all && A.traj_t) {
            CW_CK(ctx, launch_traj_times(A, cfg));
            ctx->last_launches += 1;
        }
    }
    return CW_OK;
}

static int validate(cw_ctx *ctx, const cw_opts *o, const cw_batch *B)
{
    if (!ctx || !o || !B) return CW_E_INVALID;
    if (B->n_orbits < 0 || B->n_orbits > (int64_t)1 << 31) return fail(ctx, CW_E_INVALID, "n_orbits out of range");
    if (B->n_orbits > 0 && (!B->w0 || !B->t1 || !B->t2 || !B->dt || !B->to_myr || !B->grid_flags))
        return fail(ctx, CW_E_INVALID, "null input array");
    if (o->mem_space != CW_MEM_HOST && o->mem_space != CW_MEM_DEVICE) return fail(ctx, CW_E_INVALID, "bad mem_space");
    if (o->integrator != CW_LEAPFROG && o->integrator != CW_DOP853 && o->integrator != CW_DOP853_DENSE) return fail(ctx, CW_E_UNSUPPORTED, "unknown integrator");
    if (ctx->slots.empty()) return fail(ctx, CW_E_NO_DEVICE, "context has no device");
    return CW_OK;
}

// read one int64 that may live on the device
static int fetch_i64(cw_ctx *ctx, const int64_t *p, bool on_device, int64_t *out)
{
    if (!on_device) { *out = *p; return CW_OK; }
    CW_CK(ctx, cudaMemcpy(out, p, 8, cudaMemcpyDeviceToHost));
    return CW_OK;
}

static int integrate_impl(cw_ctx *ctx, const cw_opts *o, const cw_batch *B, const cw_result *R, bool integrate,
                          int32_t *n_nodes_out, int32_t *ev_step_out)
{
    int rc = validate(ctx, o, B);
    if (rc != CW_OK) return rc;
    if (integrate && !ctx->has_pot) return fail(ctx, CW_E_NO_POTENTIAL, "cw_set_potential has not been called");
    if (integrate && (!R || !R->w_final || !R->status)) return fail(ctx, CW_E_INVALID, "w_final and status are required");
    ctx->last_launches = 0;
    for (Slot &s : ctx->slots) { s.last_kernel_ms = s.last_total_ms = 0.0; }
    const int64_t n = B->n_orbits;
    if (n == 0) return CW_OK;
    const bool dev_mode = o->mem_space == CW_MEM_DEVICE;
    const bool store_all = integrate && R->traj_off;
    if (store_all && (!R->traj_pos || !R->traj_vel)) return fail(ctx, CW_E_INVALID, "traj_pos / traj_vel missing");

    DeviceGuard guard;
    const int prev = guard.prev;
    // ---- cut the batch: by index over the device slots, then (host buffers) into pipelined chunks
    const int nslots = dev_mode ? 1 : (int)std::min<int64_t>((int64_t)ctx->slots.size(), std::max<int64_t>(1, n));
    std::vector<Shard> shards;
    for (int d = 0; d < nslots; d++) {
        const int64_t a = n * d / nslots, b = n * (d + 1) / nslots;
        // DOPRI853 runs are bound by their most expensive orbits, not by the copies (10^6 orbits: 2.4 ms of H2D
        // against 150 ms of kernel): cutting them into chunks would multiply the pilot launches and the tails
        // (measured 3.5e9 vs 6.2e9 intervals/s end to end), so they go to the device in one piece.
        const bool pipelined = !dev_mode && (!integrate || o->integrator == CW_LEAPFROG);
        int64_t Q = 0;
        if (pipelined && integrate && !store_all) {
            int prev_dev = 0;
            cudaGetDevice(&prev_dev);
            cudaSetDevice(ctx->slots[d].dev);
            Q = leapfrog_lane_capacity(ctx->pot, ctx->slots[d].sm_count);
            cudaSetDevice(prev_dev);
            // only batches of several waves: a batch of one or two (10^5 binaries with their own horizons, config 2)
            // pipelines better in the finer old chunks -- measured 1.34e11 against 0.90e11 steps/s end to end
            if (getenv("CW_NO_WAVE_CHUNKS") || b - a < 4 * Q) Q = 0;
        }
        const int64_t chunk = pipelined ? host_chunk(b - a, Q) : (b - a);
        // The first H2D copy and the last D2H copy have nothing to hide behind: ramp the chunk size up from
        // chunk/8 at the start and back down at the end (the kernels in between keep the full size).
        std::vector<int64_t> sizes;
        if (!pipelined || b - a < 4 * chunk) {
            for (int64_t c = a; c < b; c += chunk) sizes.push_back(std::min(chunk, b - c));
        } else {
            std::vector<int64_t> head, tail;
            int64_t used = 0;
            if (Q > 0) {     // whole waves here too: Q, 2Q, 4Q, ...
                for (int64_t sz = Q; sz < chunk; sz *= 2) { head.push_back(sz); used += sz; }
                for (int64_t sz = Q; sz < chunk; sz *= 2) { tail.push_back(sz); used += sz; }
            } else {
                for (int64_t sz = std::max<int64_t>(chunk / 8, HOST_CHUNK_MIN / 2); sz < chunk; sz *= 2) { head.push_back(sz); used += sz; }
                for (int64_t sz = std::max<int64_t>(chunk / 4, HOST_CHUNK_MIN / 2); sz < chunk; sz *= 2) { tail.push_back(sz); used += sz; }
            }
            int64_t mid = (b - a) - used;
            sizes = head;
            while (mid > 0) { const int64_t sz = std::min(chunk, mid); sizes.push_back(sz); mid -= sz; }
            for (size_t k = tail.size(); k-- > 0;) sizes.push_back(tail[k]);
        }
        int lane = 0;
        int64_t c = a;
        for (int64_t sz : sizes) {
            // store_all from host buffers: a shard's trajectories live in its lane's buffer, so shards are also
            // cut by trajectory COLUMNS (<= 2^26 = 3.2 GB of pos + vel); with the lane buffers re-used this
            // bounds device memory however long the orbits are
            int64_t s0 = c;
            const int64_t e0 = c + sz;
            while (s0 < e0) {
                int64_t e1 = e0;
                if (store_all && !dev_mode) {
                    const int64_t lim = R->traj_off[s0] + traj_shard_cols();
                    const int64_t *ub = std::upper_bound(R->traj_off + s0 + 1, R->traj_off + e0 + 1, lim);
                    e1 = std::min(e0, std::max<int64_t>(s0 + 1, (int64_t)(ub - R->traj_off) - 1));
                }
                Shard sh;
                sh.slot = d;
                sh.lane = dev_mode ? 0 : (lane++ % Slot::N_LANES);
                sh.a = s0;
                sh.b = e1;
                shards.push_back(sh);
                s0 = e1;
            }
            c = e0;
        }
    }
    for (Shard &sh : shards) {
        if (B->ev_off) {
            if ((rc = fetch_i64(ctx, B->ev_off + sh.a, dev_mode, &sh.ea)) != CW_OK) return rc;
            if ((rc = fetch_i64(ctx, B->ev_off + sh.b, dev_mode, &sh.eb)) != CW_OK) return rc;
        }
        if (store_all) {
            if ((rc = fetch_i64(ctx, R->traj_off + sh.a, dev_mode, &sh.ta)) != CW_OK) return rc;
            if ((rc = fetch_i64(ctx, R->traj_off + sh.b, dev_mode, &sh.tb)) != CW_OK) return rc;
        }
        if (sh.eb < sh.ea || sh.tb < sh.ta) return fail(ctx, CW_E_INVALID, "offsets must be non-decreasing");
        if (sh.eb > sh.ea && (!B->ev_time || !B->ev_dv)) return fail(ctx, CW_E_INVALID, "ev_off lists events but ev_time / ev_dv is null");
    }
    cw_result Rl;
    memset(&Rl, 0, sizeof Rl);
    if (R) Rl = *R;
    if (!integrate) { Rl.n_nodes = n_nodes_out; Rl.ev_step = ev_step_out; }

    // ---- size and reserve each slot's arena once, for all of its shards
    std::vector<Plan> plans(shards.size());
    std::vector<size_t> need((size_t)nslots * Slot::N_LANES, 0);
    for (size_t k = 0; k < shards.size(); k++) {
        size_t &nd = need[(size_t)shards[k].slot * Slot::N_LANES + shards[k].lane];
        nd = std::max(nd, shard_bytes(o, &Rl, shards[k], integrate, plans[k]));
    }
    for (int d = 0; d < nslots; d++) {
        Slot &s = ctx->slots[d];
        cudaSetDevice(s.dev);
        size_t n_sh = 0;
        for (const Shard &sh : shards) n_sh += (sh.slot == d);
        for (int l = 0; l < Slot::N_LANES; l++) {
            const size_t nd = need[(size_t)d * Slot::N_LANES + l];
            if (nd == 0) continue;
            // a stream still running a previous no_sync call must not see its buffers move
            if (nd + 4096 > s.lane_arena[l].cap) cudaStreamSynchronize(s.lanes[l]);
            CW_CK(ctx, s.lane_arena[l].reserve(nd + 4096));
        }
        if (Arena::padded(8, 8) * n_sh + 4096 > s.meta.cap)
            for (int l = 0; l < Slot::N_LANES; l++) cudaStreamSynchronize(s.lanes[l]);
        CW_CK(ctx, s.meta.reserve(Arena::padded(8, 8) * n_sh + 4096));
        CW_CK(ctx, s.ensure_events(3 * n_sh));
    }
    // ---- phase 1: enqueue everything (H2D, kernels, D2H) on every device, all asynchronous
    std::vector<size_t> ev_used(nslots, 0);
    std::vector<unsigned long long> hs(4 * shards.size(), 0);
    const int64_t total_cols = (store_all && !dev_mode) ? R->traj_off[n] : 0;
    for (size_t k = 0; k < shards.size(); k++) {
        const Shard &sh = shards[k];
        Slot &s = ctx->slots[sh.slot];
        Plan &pl = plans[k];
        cudaSetDevice(s.dev);
        pl.stream = (dev_mode && o->stream) ? (cudaStream_t)o->stream : s.lanes[sh.lane];
        pl.ev_t0 = s.events[ev_used[sh.slot]++];
        pl.ev_k0 = s.events[ev_used[sh.slot]++];
        pl.ev_k1 = s.events[ev_used[sh.slot]++];
        rc = plan_shard(ctx, s, o, B, &Rl, sh, integrate, pl);
        if (rc == CW_OK) rc = run_shard(ctx, s, o, pl, integrate, store_all);
        if (rc != CW_OK) { cudaSetDevice(prev); return rc; }
        if (!dev_mode) {
            cudaStream_t stream = pl.stream;
            const IntegArgs &A = pl.A;
            const int64_t m = sh.b - sh.a, K = sh.eb - sh.ea, Lt = sh.tb - sh.ta;
            int32_t *nn = integrate ? R->n_nodes : n_nodes_out;
            int32_t *es = integrate ? R->ev_step : ev_step_out;
            if (nn) CW_CK(ctx, cudaMemcpyAsync(nn + sh.a, A.n_nodes, m * 4, cudaMemcpyDeviceToHost, stream));
            if (es && K > 0) CW_CK(ctx, cudaMemcpyAsync(es + sh.ea, A.ev_step, K * 4, cudaMemcpyDeviceToHost, stream));
            if (integrate) {
                for (int c = 0; c < 6; c++)
                    CW_CK(ctx, cudaMemcpyAsync(R->w_final + c * n + sh.a, A.w_final + c * m, m * 8, cudaMemcpyDeviceToHost, stream));
                CW_CK(ctx, cudaMemcpyAsync(R->status + sh.a, A.status, m * 4, cudaMemcpyDeviceToHost, stream));
                if (store_all && Lt > 0) {
                    for (int c = 0; c < 3; c++) {
                        CW_CK(ctx, cudaMemcpyAsync(R->traj_pos + c * total_cols + sh.ta, A.traj_pos + c * A.traj_total, Lt * 8, cudaMemcpyDeviceToHost, stream));
                        CW_CK(ctx, cudaMemcpyAsync(R->traj_vel + c * total_cols + sh.ta, A.traj_vel + c * A.traj_total, Lt * 8, cudaMemcpyDeviceToHost, stream));
                    }
                    if (R->traj_t) CW_CK(ctx, cudaMemcpyAsync(R->traj_t + sh.ta, A.traj_t, Lt * 8, cudaMemcpyDeviceToHost, stream));
                }
            }
        }
    }
    if (dev_mode && o->no_sync) { cudaSetDevice(prev); return CW_OK; }
    // ---- phase 2: wait, gather statistics and kernel times
    int64_t stats[4] = {0, 0, 0, 0};
    for (size_t k = 0; k < shards.size(); k++) {
        Slot &s = ctx->slots[shards[k].slot];
        cudaSetDevice(s.dev);
        if (integrate && R->stats)      // (pageable destination: this copy blocks the host, so it comes after every enqueue)
            CW_CK(ctx, cudaMemcpyAsync(&hs[4 * k], plans[k].A.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, plans[k].stream));
    }
    for (size_t k = 0; k < shards.size(); k++) {
        Slot &s = ctx->slots[shards[k].slot];
        cudaSetDevice(s.dev);
        CW_CK(ctx, cudaStreamSynchronize(plans[k].stream));
        for (int q = 0; q < 4; q++) stats[q] += (int64_t)hs[4 * k + q];
        if (plans[k].timed) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, plans[k].ev_k0, plans[k].ev_k1) == cudaSuccess) s.last_kernel_ms += ms;
            if (cudaEventElapsedTime(&ms, plans[k].ev_t0, plans[k].ev_k1) == cudaSuccess) s.last_total_ms += ms;
        }
    }
    if (getenv("CW_DEBUG") && !shards.empty() && plans[0].timed) {
        // shard timeline relative to the first shard's first kernel: pre-pass start, integrator start / end
        for (size_t k = 0; k < shards.size(); k++) {
            if (shards[k].slot != 0 || !plans[k].timed) continue;
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, plans[0].ev_t0, plans[k].ev_t0);
            cudaEventElapsedTime(&b, plans[0].ev_t0, plans[k].ev_k0);
            cudaEventElapsedTime(&c, plans[0].ev_t0, plans[k].ev_k1);
            fprintf(stderr, "[cw] shard %2zu lane %d orbits %8lld: prepass %8.3f  kernel %8.3f .. %8.3f ms\n", k,
                    shards[k].lane, (long long)(shards[k].b - shards[k].a), a, b, c);
        }
    }
    if (integrate && R->stats) {
        if (dev_mode) CW_CK(ctx, cudaMemcpy(R->stats, stats, sizeof stats, cudaMemcpyHostToDevice));
        else memcpy(R->stats, stats, sizeof stats);
    }
    cudaSetDevice(prev);
    return CW_OK;
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
    const int rb = run_bytes < 0 ? -run_bytes : run_bytes;      // negative: with L2 eviction-priority hints
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
