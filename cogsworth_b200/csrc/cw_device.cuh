// cw_device.cuh -- device-side building blocks shared by every kernel of the path:
// FP64 special functions built for the DFMA pipe, the potential parameter block and the
// analytic composite gradient (the B200 replacement for gala's builtin_potentials.c +
// cpotential.c c_gradient; algorithm restated in oracle/cw_oracle.c).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cogsworth_b200.h"

namespace cw {

// ------------------------------------------------------------------------------------------
// FP64 special functions.  B200 has no FP64 SFU: sqrt / divide / log are DFMA sequences seeded
// by MUFU.RSQ64H / MUFU.RCP64H (the only FP64-related ops that run OFF the FP64 pipe).  The CUDA
// library versions carry special-case branches (denormals, inf, negative) that orbits never hit;
// these versions are branch-free and cost 5 / 3 / 12-19 FP64-pipe instructions.
// ------------------------------------------------------------------------------------------

// MUFU.RSQ64H produces only the HIGH word of the seed; `rsqrt.approx.f64` therefore costs a second
// instruction that zeroes the low word.  The refinement does not care what the low word is (any value
// perturbs the seed by < 2^-20, and the third-order steps below leave (2^-19)^3 << 2^-53), so the seed is
// assembled from the MUFU result and the low word of `donor`, a value that dies here: ptxas can then keep
// the seed in the donor's register pair and the zeroing move disappears (one issue slot per seed).
__device__ __forceinline__ double rsqrt_seed(double x, double donor)
{
#ifdef CW_NO_SEED_DONOR
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    double y;
    asm("{\n\t.reg .b32 dl, dh, yl, yh;\n\t.reg .f64 s;\n\t"
        "rsqrt.approx.ftz.f64 s, %1;\n\t"
        "mov.b64 {yl, yh}, s;\n\t"
        "mov.b64 {dl, dh}, %2;\n\t"
        "mov.b64 %0, {dl, yh};\n\t}"
        : "=d"(y) : "d"(x), "d"(donor));
    return y;
#endif
}

__device__ __forceinline__ double rcp_seed(double x, double donor)
{
#ifdef CW_NO_SEED_DONOR
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    double y;
    asm("{\n\t.reg .b32 dl, dh, yl, yh;\n\t.reg .f64 s;\n\t"
        "rcp.approx.ftz.f64 s, %1;\n\t"
        "mov.b64 {yl, yh}, s;\n\t"
        "mov.b64 {dl, dh}, %2;\n\t"
        "mov.b64 %0, {dl, yh};\n\t}"
        : "=d"(y) : "d"(x), "d"(donor));
    return y;
#endif
}

// 1/sqrt(x), x normal and positive.  Seed (~2^-22 rel. error) + one third-order step:
// y(1 + e/2 + 3e^2/8), e = 1 - x y^2  -> error O(e^3) plus rounding (<= ~1 ulp).
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);
    double p = fma(0.375, e, 0.5);
    double q = y * e;
    return fma(q, p, y);
}

// the same with the seed's low word taken from `donor` (a value that is dead after this call)
__device__ __forceinline__ double rsqrt_fast(double x, double donor)
{
    const double y = rsqrt_seed(x, donor);
    double t = x * y;
    double e = fma(-t, y, 1.0);
    double p = fma(0.375, e, 0.5);
    double q = y * e;
    return fma(q, p, y);
}

__device__ __forceinline__ double rcp_fast(double x, double donor)
{
    const double y = rcp_seed(x, donor);
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

// 1/x, x normal.  Seed + one third-order step y(1 + e + e^2), e = 1 - x y.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

// ln(x) for normal positive x (here x = 1 + r/r_s >= 1).  Classic reduction x = 2^k m,
// m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), ln m = 2s + s R(s^2) with the degree-7 minimax
// polynomial of fdlibm's e_log.c (|s| <= 0.1716, error < 2^-58.45).  Exponent handling is integer
// work on the ALU pipe, k -> double by the 2^52 bias trick (no I2F): 1 MUFU + 16 FP64-pipe
// instructions, no branches, no special cases.
__device__ __forceinline__ double log_fast(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;        // m in [1, 2)
    const int big = (hi >= 0x3ff6a09f);          // m >= ~sqrt(2): halve it
    hi -= big << 20;
    k += big;
    const double m = __hiloint2double(hi, lo);
    const double kd = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0; // 2^52 + 2^31
    const double s = (m - 1.0) * rcp_fast(m + 1.0);
    const double z = s * s;
    double R = 1.479819860511658591e-01;
    R = fma(R, z, 1.531383769920937332e-01);
    R = fma(R, z, 1.818357216161805012e-01);
    R = fma(R, z, 2.222219843214978396e-01);
    R = fma(R, z, 2.857142874366239149e-01);
    R = fma(R, z, 3.999999999940941908e-01);
    R = fma(R, z, 6.666666666666735130e-01);
    const double r = fma(s, R * z, s + s);       // 2s + s z R
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    return fma(kd, LN2_HI, fma(kd, LN2_LO, r));
}

// ------------------------------------------------------------------------------------------
// Potential parameter block.  Passed by value as a __grid_constant__ kernel parameter, i.e. it
// lives in the constant bank (c[0x0][...]) and every access is a uniform broadcast -- the
// per-launch equivalent of a __constant__ symbol, but safe with several devices/streams per
// process.
// ------------------------------------------------------------------------------------------
enum PotKind : int {
    POT_GENERIC = 0, // any ordered list of components (loop + switch)
    POT_MW_V1 = 1,   // MiyamotoNagai + Hernquist + Hernquist + spherical NFW
    POT_MW_V2 = 2,   // 3 x MiyamotoNagai sharing b (MN3 disk) + Hernquist + Hernquist + NFW
    POT_GENERAL = 3, // full gradient vector: rotated / non-diagonal components, time-interpolated amplitudes
    // Separable time dependence Phi(q, t) = f(t) Phi_0(q): every mass follows the same curve (the reference's
    // evolving-potential tutorial: one mass-fraction array for all components).  The static kernels with one spline
    // evaluation and three multiplications per gradient; f(t) is stored as the knots of "component 0".
    POT_MW_V2_T = 4,
    POT_GENERIC_T = 5,
    // The static, diagonal subset of POT_GENERAL: no time-interpolated amplitude and no non-diagonal component, only
    // rotated frames (LM10: disk and bulge beside a rotated triaxial logarithmic halo).  The same arithmetic as
    // POT_GENERAL bit for bit (a static amplitude is the factor 1.0), without the spline look-up and the called
    // non-diagonal function in the kernel: no spills, a third fewer instructions per step.
    POT_ROTATED = 6,
    // gala's LM10Potential as a compile-time shape of POT_ROTATED: Miyamoto-Nagai disk, Hernquist bulge, triaxial
    // logarithmic halo in a rotated frame.  The same source expressions with the component list known to the compiler:
    // no type dispatch, no per-lane constant-bank indices (241 -> ~130 warp instructions per step).
    POT_LM10 = 7
};
template <int KIND> struct KindTraits {
    static constexpr bool TS = (KIND == POT_MW_V2_T || KIND == POT_GENERIC_T);
    static constexpr int BASE = (KIND == POT_MW_V2_T) ? POT_MW_V2 : (KIND == POT_GENERIC_T) ? POT_GENERIC : KIND;
    // the force functions return the gradient vector itself, not three diagonal factors
    static constexpr bool VEC = (KIND == POT_GENERAL || KIND == POT_ROTATED || KIND == POT_LM10);
};

struct PotParams {
    int32_t kind;
    int32_t n_comp;
    double G;
    int32_t type[CW_MAX_COMPONENTS];
    double GM[CW_MAX_COMPONENTS]; // fl(G * m), the same product gala forms on every call
    double p1[CW_MAX_COMPONENTS]; // a | c | r_s
    double p2[CW_MAX_COMPONENTS]; // b | - | 1/r_s
    double q2[CW_MAX_COMPONENTS][3]; // triaxial NFW: 1/a^2, 1/b^2, 1/c^2
    // pre-digested copies for the specialised kinds (MN terms first, then the spherical ones)
    double mn_GM[3], mn_a[3], mn_b2[3];
    double mn_GMa[3];             // fl(GM_i a_i): the vertical sum of the shared-b disk accumulates GM_i a_i D_i^-3/2
    double h_GM[2], h_c[2];
    double nfw_GM, nfw_rs, nfw_inv_rs;
    // Full 64-bit numeric constants of log_tab.  As kernel parameters they are constant-bank operands
    // of the DFMA that uses them; as literals ptxas rebuilds each with two UMOVs inside the hot loop,
    // and on sm_100 every non-FP64 instruction costs half an FP64 issue slot (an FP64 warp instruction
    // holds the dispatch port for 2 cycles; measured, tools/exp/pipe_mix2.cu).
    double k_ln2, k_c5, k_c3, k_c158;
    // ---- POT_GENERAL only (cw_set_potential_general)
    int32_t has_R[CW_MAX_COMPONENTS];
    double R[CW_MAX_COMPONENTS][9];        // row-major: q' = R q, grad = R^T grad'
    double pe[CW_MAX_COMPONENTS][4];       // pre-digested parameters of the non-diagonal components
    int32_t k_off[CW_MAX_COMPONENTS + 1];  // knots of component i: [k_off[i], k_off[i+1]); empty = static
    double k_t[CW_MAX_KNOTS];              // knot times [Myr]
    double k_c[CW_MAX_KNOTS][4];           // amplitude relative to GM[i] on [k_t[j], k_t[j+1]]:
                                           // c0 + s (c1 + s (c2 + s c3)), s = t - k_t[j]
    int32_t run_end[CW_MAX_COMPONENTS];    // > i: components [i, run_end[i]) are static, unrotated and diagonal: one
                                           // force_generic_range call that shares the radius between them
    double k_t0[CW_MAX_COMPONENTS];        // equally spaced knots (np.linspace): first knot and 1 / spacing, else 0
    double k_inv_dt[CW_MAX_COMPONENTS];
    double t_eval;                         // time of the pointwise kernels
};

__host__ __device__ inline void pot_fill_constants(PotParams &P)
{
    P.k_ln2 = 6.931471805599453094e-01;
    P.k_c5 = 0.2;
    P.k_c3 = 1.0 / 3.0;
    P.k_c158 = 1.875;                   // 15/8: second-order coefficient of (1 - e)^-3/2
}

// Table-driven ln(x), x >= 1 (Tang-style): x = 2^k m, m in [1,2); c_j = 1 + j/1024 is the table node
// nearest to m (j = 0..1024), r = m/c_j - 1 (one FMA with the tabulated 1/c_j, |r| <= 2^-11) and
// ln m = ln c_j + log1p(r), log1p by its degree-5 Taylor polynomial (remainder r^6/6 < 2^-57 |r|).
// The 1025 x {1/c_j, -ln(1/c_j)} table (16 KB) sits in shared memory and is filled once per CTA by
// log_table_init; the look-up is one LDS.128 on the otherwise idle LSU pipe, addressed through the
// CTA's 32-bit shared window address (no uniform-register address arithmetic inside the hot loop).
// 1 DADD for k plus 9 FP64-pipe instructions, no MUFU: 9 fewer FP64 instructions than log_fast.
// (Round 1 used 128 nodes and degree 7: two more DFMA and two more 64-bit constants per call; every
// instruction of the leapfrog loop costs issue slots -- FP64 two, anything else one -- and the kernel
// is exactly issue-bound, so both the DFMAs and the constant loads showed.)
constexpr int LOG_TABLE_BITS = 10;
constexpr int LOG_TABLE_N = (1 << LOG_TABLE_BITS) + 1;

struct LogTab {
    uint32_t s;      // shared-window byte address of the table, 0xffffffff = no table (use log_fast)
    // Seed donors (see rsqrt_seed) for the two reciprocal square roots of the Milky Way gradient whose arguments stay
    // alive: two values of the CALLER that die at this gradient evaluation (the leapfrog passes the scaled force factors
    // of the previous step).  Only their low words are read, and no result depends on them.
    double d0, d1;
    __device__ __forceinline__ bool on() const { return s != 0xffffffffu; }
};
__device__ __forceinline__ LogTab no_log_table() { LogTab t; t.s = 0xffffffffu; t.d0 = t.d1 = 0.0; return t; }

__device__ __forceinline__ LogTab log_table_init(double2 *tab)
{
    for (int j = threadIdx.x; j < LOG_TABLE_N; j += blockDim.x) {
        const double invc = 1.0 / (1.0 + (double)j * (1.0 / (double)(1 << LOG_TABLE_BITS)));
        tab[j] = make_double2(invc, -log(invc));      // ln of the c actually used (1/invc)
    }
    __syncthreads();
    LogTab t;
    t.s = (uint32_t)__cvta_generic_to_shared(tab);
    t.d0 = t.d1 = 0.0;
    return t;
}

__device__ __forceinline__ double log_tab(const PotParams &P, double x, LogTab tab)
{
    const int hi = __double2hiint(x);
    const int k = (hi >> 20) - 1023;
    // byte offset 16 j of node j = round(top LOG_TABLE_BITS+1 mantissa bits / 2): (((hi >> 6) & 0x3ff8) + 8) & ~15
    const uint32_t addr = ((((uint32_t)hi >> (16 - LOG_TABLE_BITS)) & ((2u << (LOG_TABLE_BITS + 3)) - 8u)) + 8u + tab.s) & ~15u;
    const double kd = (double)k;                                     // I2F on the XU pipe: one issue slot
    double tx, ty;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(addr));
    // r = m / c_j - 1 with m = x 2^-k: the power of two goes into the exponent of the tabulated 1/c_j (exact),
    // so the mantissa of x never has to be re-assembled into a register pair of its own
    const double txs = __hiloint2double(__double2hiint(tx) - (hi & 0x7ff00000) + 0x3ff00000, __double2loint(tx));
    const double r = fma(x, txs, -1.0);
    double q = fma(P.k_c5, r, -0.25);
    q = fma(q, r, P.k_c3);
    q = fma(q, r, -0.5);
    const double lp = fma(r * r, q, r);                              // log1p(r)
    // k <= 1023 and the result is compared with u / (1 + u) right away: one correctly rounded ln 2 is enough
    // (k (LN2 - ln 2) < 6e-17 k, relative to k ln 2 + ... below 1e-16); measured by cw_selftest_math
    return fma(kd, P.k_ln2, ty + lp);
}

// Specialised composite: NMN Miyamoto-Nagai terms (+ optionally one shared b), two Hernquist
// spheres and one spherical NFW halo.  FP64-pipe instruction budget (v2): ~105 per gradient.
//   MN(m,a,b):   sz = sqrt(z^2+b^2); zd = a+sz; f = GM (R^2+zd^2)^-3/2;  g += (f x, f y, f z (1 + a/sz))
//   Hernquist:   f = GM / ((r+c)^2 r);                                    g += f q
//   NFW:         f = GM (ln(1+r/rs) - r/(r+rs)) / r^3;                    g += f q
template <int NMN, bool SHARED_B>
__device__ __forceinline__ void force_mw(const PotParams &P, LogTab ltab,
                                         double x, double y, double z, double &Fxy, double &Fz)
{
    const double R2 = fma(y, y, x * x);
    const double z2 = z * z;
    const double r2 = R2 + z2;

    // ---- the three spherical components over ONE reciprocal square root and ONE reciprocal
    const double ir = rsqrt_fast(r2, ltab.d0);             // seed donors: values whose last use is here
    const double r = r2 * ir;
    // two Hernquist spheres over one reciprocal: GM0/rc0^2 + GM1/rc1^2 = (GM0 rc1^2 + GM1 rc0^2) / (rc0 rc1)^2
    const double rc0 = r + P.h_c[0], rc1 = r + P.h_c[1];
    const double s0 = rc0 * rc0, s1 = rc1 * rc1;
    // NFW
    const double u = r * P.nfw_inv_rs;
    const double w = 1.0 + u;
    // ... and over the same reciprocal 1 / (s0 s1 w): 1/(s0 s1) = inv w, 1/w = inv s0 s1
    const double s01 = s0 * s1;
    const double inv = rcp_fast(s01 * w, rc0);
    const double h = fma(P.h_GM[1], s0, P.h_GM[0] * s1) * (inv * w);
    const double lw = ltab.on() ? log_tab(P, w, ltab) : log_fast(w);
    const double bracket = lw - u * (inv * s01);
    // fs = h / r + GM_nfw bracket / r^3 = (h + (GM_nfw / r^2) bracket) / r
    const double fs = ir * fma(P.nfw_GM * (ir * ir), bracket, h);

    // ---- Miyamoto-Nagai terms.  One term needs f = GM D^-3/2, not D^-1/2 itself, so the third-order correction is
    // applied to the cube directly: with the seed y ~ D^-1/2 and e = 1 - D y^2,
    //     D^-3/2 = y^3 (1 - e)^-3/2 = y^3 (1 + e (3/2 + 15/8 e) + O(e^3)),   |e| < 2^-21,
    // and y^2 serves both e and y^3.
    if (SHARED_B) {
        // all terms share sqrt(z^2 + b^2): fz = sum f_i (1 + a_i/sz) = fxy + (sum a_i f_i) / sz.  The masses enter only
        // in the two accumulations (GM_i and GM_i a_i are constant-bank operands of their FMAs), the first of which
        // starts from the spherical part: 8 FP64 instructions per term + 6 for the sums (round 1: 9 + 6).
        const double q = z2 + P.mn_b2[0];
        const double isz = rsqrt_fast(q, ltab.d1);
        const double sz = q * isz;
        double t[NMN];
#pragma unroll
        for (int i = 0; i < NMN; i++) {
            const double zd = P.mn_a[i] + sz;
            const double D = fma(zd, zd, R2);
            const double y = rsqrt_seed(D, zd);            // zd dies here
            const double y2 = y * y;
            const double e = fma(-D, y2, 1.0);
            const double p = fma(P.k_c158, e, 1.5);
            const double y3 = y2 * y;
            t[i] = fma(y3 * e, p, y3);
        }
        double fxy = fs, fa = P.mn_GMa[0] * t[0];
#pragma unroll
        for (int i = 0; i < NMN; i++) {
            fxy = fma(P.mn_GM[i], t[i], fxy);
            if (i > 0) fa = fma(P.mn_GMa[i], t[i], fa);
        }
        Fxy = fxy;                   // grad = (Fxy x, Fxy y, Fz z)
        Fz = fma(isz, fa, fxy);
    } else {
        double fxy = fs, za = 0.0;
#pragma unroll
        for (int i = 0; i < NMN; i++) {
            const double q = z2 + P.mn_b2[i];
            const double isz = rsqrt_fast(q);
            const double zd = fma(q, isz, P.mn_a[i]);
            const double rs = rsqrt_fast(fma(zd, zd, R2));
            const double f = (rs * rs) * (P.mn_GM[i] * rs);
            fxy += f;
            const double fa = P.mn_a[i] * f;                 // f_i a_i / sz_i accumulates in za
            za = (i == 0) ? fa * isz : fma(fa, isz, za);
        }
        Fxy = fxy;
        Fz = fxy + za;
    }
}

// ------------------------------------------------------------------------------------------
// POT_GENERAL: every component evaluated on its own, in its own (rotated) frame, with its own clock.
// Precision before speed here: IEEE sqrt / divide and the CUDA math library.
// ------------------------------------------------------------------------------------------

// relative amplitude m(t) / m_ref of a gp.TimeInterpolatedPotential component (1 for a static one)
__device__ __forceinline__ double amplitude_rel(const PotParams &P, int i, double t)
{
    const int a = P.k_off[i], b = P.k_off[i + 1];
    if (b <= a) return 1.0;
    if (b - a == 1) return P.k_c[a][0];
    t = fmin(fmax(t, P.k_t[a]), P.k_t[b - 1]);     // the end values are held outside the knots
    int lo = a, hi = b - 2;                        // interval index: k_t[j] <= t <= k_t[j+1]
    if (P.k_inv_dt[i] > 0.0) {
        // equally spaced knots: the index directly.  Rounding may pick the neighbour of the exact interval when t
        // sits within an ulp of a knot; the two cubics agree there.
        lo = a + min(max(__double2int_rd((t - P.k_t0[i]) * P.k_inv_dt[i]), 0), hi - a);
    } else {
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (t >= P.k_t[mid]) lo = mid; else hi = mid - 1;
        }
    }
    const double s = t - P.k_t[lo];
    return fma(s, fma(s, fma(s, P.k_c[lo][3], P.k_c[lo][2]), P.k_c[lo][1]), P.k_c[lo][0]);
}

// regularised lower incomplete gamma function P(a, x), a > 0, x >= 0: power series for x < a + 1,
// Lentz continued fraction of Q = 1 - P otherwise (converged to 1 ulp-ish; <= 300 terms)
__host__ __device__ inline double gamma_p(double a, double x, double lgam_a)
{
    if (!(x > 0.0)) return 0.0;
    const double pref = exp(a * log(x) - x - lgam_a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 500; n++) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * pref;
    }
    const double tiny = 1e-300;
    double bb = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / bb, h = d;
    for (int n = 1; n < 500; n++) {
        const double an = -(double)n * ((double)n - a);
        bb += 2.0;
        d = an * d + bb;
        if (fabs(d) < tiny) d = tiny;
        c = bb + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - pref * h;
}

// gradient (and, when want_phi, value) of ONE non-diagonal component in its own frame; amp = GM[i] * rel.
// Deliberately NOT inlined, with every parameter passed by value: the formulas below leave FMA contraction to the
// compiler, and one compiled copy is what makes the final-state kernels return the bits of the store_all kernels
// (an inlined copy per kernel was contracted differently in each -- caught by the bit-equality test).
struct Vec4 { double x, y, z, w; };
static __device__ __noinline__ Vec4 nondiag_component_v(int t, double amp, double pe0, double pe1, double pe2, double pe3,
                                                        double x, double y, double z, bool want_phi)
{
    Vec4 o;
    o.w = 0.0;
    if (t == CW_COMP_LONG_MURALI_BAR) {          // pe = (a, b, c^2)
        const double a = pe0;
        const double sz = sqrt(fma(z, z, pe2));
        const double B = pe1 + sz;
        const double D = fma(y, y, B * B);
        const double xm = a - x, xp = a + x;
        const double Tm = sqrt(fma(xm, xm, D)), Tp = sqrt(fma(xp, xp, D));
        const double sT = Tm + Tp;
        const double f1 = amp / (2.0 * Tm * Tp);
        o.x = 4.0 * f1 * x / sT;
        const double cm = f1 / D * (sT - 4.0 * x * x / sT);
        o.y = cm * y;
        o.z = cm * B * z / sz;
        if (want_phi) o.w = amp / (2.0 * a) * log((x - a + Tm) / (x + a + Tp));
    } else if (t == CW_COMP_LEESUTO_NFW) {       // amp = phi0, pe = (1 / r_s, e_b^2, e_c^2)
        const double eb2 = pe1, ec2 = pe2;
        const double r2 = fma(z, z, fma(y, y, x * x));
        const double r = sqrt(r2);
        const double u = r * pe0;
        const double L = log(1.0 + u);
        const double u2 = u * u, u3 = u2 * u, u4 = u2 * u2, opu = 1.0 + u;
        const double N = u2 - 3.0 * u - 6.0, Dn = 2.0 * u2 * opu;
        const double F3 = N / Dn + 3.0 * L / u3;
        const double dF1 = L / u2 - 1.0 / (u * opu);
        const double dF2 = 0.5 / u2 - 2.0 / u3 + (3.0 / u4 - 1.0 / u2) * L + (1.0 / u - 1.0 / u3) / opu;
        const double dF3 = ((2.0 * u - 3.0) * Dn - N * (4.0 * u + 6.0 * u2)) / (Dn * Dn) + 3.0 / (u3 * opu) - 9.0 * L / u4;
        const double E = 0.5 * (eb2 + ec2);
        const double S = 0.5 * (eb2 * y * y + ec2 * z * z);
        const double radial = (dF1 + E * dF2 + S / r2 * dF3) * pe0 / r;
        const double c3 = F3 / r2;
        const double w = 2.0 * S / r2;
        o.x = amp * (radial * x - c3 * w * x);
        o.y = amp * (radial * y + c3 * (eb2 - w) * y);
        o.z = amp * (radial * z + c3 * (ec2 - w) * z);
        if (want_phi) {
            const double F1 = -L / u;
            const double F2 = -1.0 / 3.0 + (2.0 * u2 - 3.0 * u + 6.0) / (6.0 * u2) + (1.0 / u - 1.0 / u3) * L;
            o.w = amp * (F1 + E * F2 + S / r2 * F3);
        }
    } else {                                     // CW_COMP_POWERLAW_CUTOFF: pe = (s = 1.5 - alpha/2, 1/r_c^2, lgamma(s), lgamma(s - 0.5))
        const double r2 = fma(z, z, fma(y, y, x * x));
        const double r = sqrt(r2);
        const double xx = r2 * pe1;
        const double Pm = gamma_p(pe0, xx, pe2);      // M(<r) / m
        const double f = amp * Pm / (r2 * r);
        o.x = f * x; o.y = f * y; o.z = f * z;
        if (want_phi) {
            // -G M(<r)/r - G m Gamma(s - 1/2, x) / (r_c Gamma(s)): the shells outside r  (needs alpha < 2)
            const double Q2 = 1.0 - gamma_p(pe0 - 0.5, xx, pe3);
            o.w = -amp * (Pm / r + Q2 * exp(pe3 - pe2) * sqrt(pe1));
        }
    }
    return o;
}

__device__ __forceinline__ void nondiag_component(const PotParams &P, int i, double amp, double x, double y, double z,
                                                  double &gx, double &gy, double &gz, double *phi)
{
    const Vec4 o = nondiag_component_v(P.type[i], amp, P.pe[i][0], P.pe[i][1], P.pe[i][2], P.pe[i][3], x, y, z,
                                       phi != nullptr);
    gx = o.x; gy = o.y; gz = o.z;
    if (phi) *phi = o.w;
}

// The component list as the force functions see it: read from the parameter block (any potential), or fixed at compile
// time (a named composite: loops unroll, type switches fold, every parameter becomes a uniform operand).
struct ShapeDyn {
    static constexpr bool FIXED = false;
    static __device__ __forceinline__ int n(const PotParams &P) { return P.n_comp; }
    static __device__ __forceinline__ int type(const PotParams &P, int i) { return P.type[i]; }
    static __device__ __forceinline__ bool rotated(const PotParams &P, int i) { return P.has_R[i] != 0; }
    static __device__ __forceinline__ int run_end(const PotParams &P, int i) { return P.run_end[i]; }
};
struct ShapeLM10 {
    static constexpr bool FIXED = true;
    static __device__ __forceinline__ int n(const PotParams &) { return 3; }
    static __device__ __forceinline__ int type(const PotParams &, int i)
    {
        return i == 0 ? CW_COMP_MIYAMOTO_NAGAI : i == 1 ? CW_COMP_HERNQUIST : CW_COMP_LOG_TRIAXIAL;
    }
    static __device__ __forceinline__ bool rotated(const PotParams &, int i) { return i == 2; }
    static __device__ __forceinline__ int run_end(const PotParams &, int i) { return i < 2 ? 2 : 0; }
};

__device__ __forceinline__ bool comp_is_nondiag(int t) { return t >= CW_COMP_LONG_MURALI_BAR; }

// Generic composite: components in the caller's order.
template <class SHAPE = ShapeDyn>
__device__ __forceinline__ void force_generic_range(const PotParams &P, LogTab ltab, const int i0, const int i1,
                                                    double x, double y, double z, double &Fxy, double &Fy, double &Fz)
{
    const double R2 = fma(y, y, x * x);
    const double z2 = z * z;
    const double r2 = R2 + z2;
    double fxy = 0.0, fz = 0.0;
    double fyx = 0.0;             // Fy - Fxy: non-zero only for components flattened differently in x and y
    double ir = 0.0, r = 0.0;
    bool have_r = false;
#pragma unroll
    for (int i = i0; i < i1; i++) {
        const int t = SHAPE::type(P, i);
        if (t == CW_COMP_MIYAMOTO_NAGAI) {
            const double q = fma(P.p2[i], P.p2[i], z2);
            const double isz = rsqrt_fast(q);
            const double zd = P.p1[i] + q * isz;
            const double rs = rsqrt_fast(fma(zd, zd, R2));
            const double f = (rs * rs) * (P.GM[i] * rs);
            fxy += f;
            fz = fma(f, fma(P.p1[i], isz, 1.0), fz);
        } else {
            if (!have_r) {
                ir = rsqrt_fast(r2);
                r = __dmul_rn(r2, ir);       // a rounded product of its own in every instantiation (with a compile-time
                                             // component list r + c would otherwise contract into fma(r2, ir, c))
                have_r = true;
            }
            double f;
            if (t == CW_COMP_LOGARITHMIC) {                 // GM = v_c^2, p1 = r_h^2, p2 = 1/q3^2
                const double inv = P.GM[i] * rcp_fast(fma(z2, P.p2[i], R2 + P.p1[i]));
                fxy += inv;
                fz = fma(inv, P.p2[i], fz);
                continue;
            }
            if (t == CW_COMP_NFW_TRIAXIAL) {                // gradient = fac (x / a^2, y / b^2, z / c^2), flattened radius
                const double m2 = fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], (x * x) * P.q2[i][0]));
                const double im = rsqrt_fast(m2);
                const double u = (m2 * im) * P.p2[i];
                const double w = 1.0 + u;
                const double bracket = (ltab.on() ? log_tab(P, w, ltab) : log_fast(w)) - u * rcp_fast(w);
                const double fac = P.GM[i] * ((im * im) * im) * bracket;
                fxy = fma(fac, P.q2[i][0], fxy);
                fyx = fma(fac, P.q2[i][1] - P.q2[i][0], fyx);
                fz = fma(fac, P.q2[i][2], fz);
                continue;
            }
            if (t == CW_COMP_LOG_TRIAXIAL) {                // GM = v_c^2, p1 = r_h^2, q2 = 1 / q_i^2
                const double S = fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], fma(x * x, P.q2[i][0], P.p1[i])));
                const double fac = P.GM[i] * rcp_fast(S);
                fxy = fma(fac, P.q2[i][0], fxy);
                fyx = fma(fac, P.q2[i][1] - P.q2[i][0], fyx);
                fz = fma(fac, P.q2[i][2], fz);
                continue;
            }
            if (t == CW_COMP_KUZMIN) {                      // p1 = a; gz = f (a + |z|) sign z
                const double az = P.p1[i] + fabs(z);
                const double is = rsqrt_fast(fma(az, az, R2));
                const double fk = P.GM[i] * ((is * is) * is);
                fxy += fk;
                if (z != 0.0) fz = fma(fk, az * rcp_fast(fabs(z)), fz);
                continue;
            }
            if (t == CW_COMP_HERNQUIST) {
                const double rc = r + P.p1[i];
                f = P.GM[i] * rcp_fast(rc * rc) * ir;
            } else if (t == CW_COMP_BURKERT) {              // GM = pi G rho r0, p1 = r0, p2 = 1 / r0
                const double xb = r * P.p2[i];
                const double br = log(fma(xb, xb, 1.0)) + 2.0 * log(1.0 + xb) - 2.0 * atan(xb);
                f = P.GM[i] * br * rcp_fast(xb * xb) * ir;
            } else if (t == CW_COMP_KEPLER) {
                f = P.GM[i] * ((ir * ir) * ir);
            } else if (t == CW_COMP_PLUMMER) {              // p2 = b^2
                const double is = rsqrt_fast(r2 + P.p2[i]);
                f = P.GM[i] * ((is * is) * is);
            } else if (t == CW_COMP_ISOCHRONE) {            // p1 = b, p2 = b^2
                const double s2 = r2 + P.p2[i];
                const double is = rsqrt_fast(s2);
                const double sb = fma(s2, is, P.p1[i]);
                f = P.GM[i] * is * rcp_fast(sb * sb);
            } else if (t == CW_COMP_JAFFE) {
                f = P.GM[i] * (ir * ir) * rcp_fast(r + P.p1[i]);
            } else { // CW_COMP_NFW_SPHERICAL
                const double u = r * P.p2[i];               // p2 = 1/r_s for NFW
                const double w = 1.0 + u;
                const double bracket = (ltab.on() ? log_tab(P, w, ltab) : log_fast(w)) - u * rcp_fast(w);
                f = P.GM[i] * ((ir * ir) * ir) * bracket;
            }
            // (explicit roundings: where the component list is known at compile time f is a visible product, and a
            // contraction into these sums would change the bits against the run-time list)
            fxy = __dadd_rn(fxy, f);
            fz = __dadd_rn(fz, f);
        }
    }
    Fxy = fxy;
    Fy = fxy + fyx;
    Fz = fz;
}

__device__ __forceinline__ void force_generic(const PotParams &P, LogTab ltab,
                                              double x, double y, double z, double &Fxy, double &Fy, double &Fz)
{
    force_generic_range(P, ltab, 0, P.n_comp, x, y, z, Fxy, Fy, Fz);
}

// The three diagonal factors of ONE component (types 1..12) for the general kernels: grad = (Fx x, Fy y, Fz z).
// The same formulas as force_generic_range, behind a jump table instead of a chain of comparisons and without the
// bookkeeping that lets a whole composite share one radius.
template <class SHAPE = ShapeDyn>
__device__ __forceinline__ void diag_component(const PotParams &P, LogTab ltab, const int i,
                                               double x, double y, double z, double &Fx, double &Fy, double &Fz)
{
    const double R2 = fma(y, y, x * x);
    const double z2 = z * z;
    const double r2 = R2 + z2;
    const double GM = P.GM[i], p1 = P.p1[i], p2 = P.p2[i];
    double f;
    switch (SHAPE::type(P, i)) {
    case CW_COMP_MIYAMOTO_NAGAI: {
        const double q = fma(p2, p2, z2);
        const double isz = rsqrt_fast(q);
        const double zd = p1 + q * isz;
        const double rs = rsqrt_fast(fma(zd, zd, R2));
        f = (rs * rs) * (GM * rs);
        Fx = Fy = f;
        Fz = f * fma(p1, isz, 1.0);
        return;
    }
    case CW_COMP_LOGARITHMIC: {                 // GM = v_c^2, p1 = r_h^2, p2 = 1/q3^2
        f = GM * rcp_fast(fma(z2, p2, R2 + p1));
        Fx = Fy = f;
        Fz = f * p2;
        return;
    }
    case CW_COMP_NFW_TRIAXIAL: {
        const double m2 = fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], (x * x) * P.q2[i][0]));
        const double im = rsqrt_fast(m2);
        const double u = (m2 * im) * p2;
        const double w = 1.0 + u;
        const double bracket = (ltab.on() ? log_tab(P, w, ltab) : log_fast(w)) - u * rcp_fast(w);
        f = GM * ((im * im) * im) * bracket;
        Fx = f * P.q2[i][0]; Fy = f * P.q2[i][1]; Fz = f * P.q2[i][2];
        return;
    }
    case CW_COMP_LOG_TRIAXIAL: {                // GM = v_c^2, p1 = r_h^2, q2 = 1 / q_i^2
        const double S = fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], fma(x * x, P.q2[i][0], p1)));
        f = GM * rcp_fast(S);
        Fx = f * P.q2[i][0]; Fy = f * P.q2[i][1]; Fz = f * P.q2[i][2];
        return;
    }
    case CW_COMP_KUZMIN: {
        const double az = p1 + fabs(z);
        const double is = rsqrt_fast(fma(az, az, R2));
        f = GM * ((is * is) * is);
        Fx = Fy = f;
        Fz = (z != 0.0) ? f * (az * rcp_fast(fabs(z))) : 0.0;
        return;
    }
    case CW_COMP_PLUMMER: {                     // p2 = b^2
        const double is = rsqrt_fast(r2 + p2);
        f = GM * ((is * is) * is);
        break;
    }
    case CW_COMP_ISOCHRONE: {                   // p1 = b, p2 = b^2
        const double s2 = r2 + p2;
        const double is = rsqrt_fast(s2);
        const double sb = fma(s2, is, p1);
        f = GM * is * rcp_fast(sb * sb);
        break;
    }
    default: {
        const double ir = rsqrt_fast(r2);
        const double r = r2 * ir;
        switch (SHAPE::type(P, i)) {
        case CW_COMP_HERNQUIST: {
            const double rc = r + p1;
            f = GM * rcp_fast(rc * rc) * ir;
            break;
        }
        case CW_COMP_BURKERT: {                 // GM = pi G rho r0, p1 = r0, p2 = 1 / r0
            const double xb = r * p2;
            const double br = log(fma(xb, xb, 1.0)) + 2.0 * log(1.0 + xb) - 2.0 * atan(xb);
            f = GM * br * rcp_fast(xb * xb) * ir;
            break;
        }
        case CW_COMP_KEPLER: f = GM * ((ir * ir) * ir); break;
        case CW_COMP_JAFFE: f = GM * (ir * ir) * rcp_fast(r + p1); break;
        default: {                              // CW_COMP_NFW_SPHERICAL: p2 = 1 / r_s
            const double u = r * p2;
            const double w = 1.0 + u;
            const double bracket = (ltab.on() ? log_tab(P, w, ltab) : log_fast(w)) - u * rcp_fast(w);
            f = GM * ((ir * ir) * ir) * bracket;
        }
        }
    }
    }
    Fx = Fy = Fz = f;
}

// POT_GENERAL: gradient at time t [Myr], components accumulated in the caller's order (c_gradient)
template <bool STATIC_DIAG, class SHAPE = ShapeDyn>
static __device__ __forceinline__ void gradient_general(const PotParams &P, LogTab ltab, double t,
                                              double x, double y, double z, double &gx, double &gy, double &gz)
{
    double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
    for (int i = 0; i < SHAPE::n(P); i++) {
        const int e = SHAPE::run_end(P, i);
        if (e > i) {       // a run of plain components (LM10's disk and bulge beside its rotated halo): evaluated as one,
                           // on its first member
            if (i == 0 || SHAPE::run_end(P, i - 1) != e) {
                double Fx, Fy, Fz;
                force_generic_range<SHAPE>(P, ltab, i, e, x, y, z, Fx, Fy, Fz);
                ax = __dadd_rn(ax, __dmul_rn(Fx, x)); ay = __dadd_rn(ay, __dmul_rn(Fy, y)); az = __dadd_rn(az, __dmul_rn(Fz, z));
            }
            continue;
        }
        const double rel = STATIC_DIAG ? 1.0 : amplitude_rel(P, i, t);
        const bool rot = SHAPE::rotated(P, i);
        double xr = x, yr = y, zr = z;
        if (rot) {
            xr = fma(P.R[i][2], z, fma(P.R[i][1], y, __dmul_rn(P.R[i][0], x)));
            yr = fma(P.R[i][5], z, fma(P.R[i][4], y, __dmul_rn(P.R[i][3], x)));
            zr = fma(P.R[i][8], z, fma(P.R[i][7], y, __dmul_rn(P.R[i][6], x)));
        }
        double cx, cy, cz;
        if (!STATIC_DIAG && comp_is_nondiag(SHAPE::type(P, i))) {
            nondiag_component(P, i, __dmul_rn(P.GM[i], rel), xr, yr, zr, cx, cy, cz, nullptr);
        } else {
            double Fx, Fy, Fz;
            diag_component<SHAPE>(P, ltab, i, xr, yr, zr, Fx, Fy, Fz);
            // (a static amplitude is the factor 1.0: the product it is left out of is exact)
            cx = __dmul_rn(STATIC_DIAG ? Fx : __dmul_rn(Fx, rel), xr);
            cy = __dmul_rn(STATIC_DIAG ? Fy : __dmul_rn(Fy, rel), yr);
            cz = __dmul_rn(STATIC_DIAG ? Fz : __dmul_rn(Fz, rel), zr);
        }
        if (rot) {
            ax = __dadd_rn(ax, fma(P.R[i][6], cz, fma(P.R[i][3], cy, __dmul_rn(P.R[i][0], cx))));
            ay = __dadd_rn(ay, fma(P.R[i][7], cz, fma(P.R[i][4], cy, __dmul_rn(P.R[i][1], cx))));
            az = __dadd_rn(az, fma(P.R[i][8], cz, fma(P.R[i][5], cy, __dmul_rn(P.R[i][2], cx))));
        } else {
            ax = __dadd_rn(ax, cx); ay = __dadd_rn(ay, cy); az = __dadd_rn(az, cz);
        }
    }
    gx = ax; gy = ay; gz = az;
}

// the vector-gradient kinds
template <int KIND>
static __device__ __forceinline__ void gradient_vec(const PotParams &P, LogTab ltab, double t,
                                                    double x, double y, double z, double &gx, double &gy, double &gz)
{
    if (KIND == POT_LM10) gradient_general<true, ShapeLM10>(P, ltab, t, x, y, z, gx, gy, gz);
    else gradient_general<KIND == POT_ROTATED, ShapeDyn>(P, ltab, t, x, y, z, gx, gy, gz);
}

// ltab: the CTA's shared-memory log table (log_table_init), or nullptr to use the table-free log.
// force<KIND> returns the two scalar factors of the axisymmetric composite: grad = (Fxy x, Fxy y, Fz z).
template <int KIND>
__device__ __forceinline__ void force(const PotParams &P, LogTab ltab,
                                      double x, double y, double z, double &Fxy, double &Fz)
{
    if (KIND == POT_MW_V1) force_mw<1, false>(P, ltab, x, y, z, Fxy, Fz);
    else if (KIND == POT_MW_V2) force_mw<3, true>(P, ltab, x, y, z, Fxy, Fz);
    else { double Fy; force_generic(P, ltab, x, y, z, Fxy, Fy, Fz); }
}

// The same with the y factor separate: grad = (Fx x, Fy y, Fz z).  Fy = Fx for the axisymmetric Milky Way kinds; the
// generic composite may contain components flattened along x or y (triaxial NFW).
template <int KIND>
__device__ __forceinline__ void force3(const PotParams &P, LogTab ltab,
                                       double x, double y, double z, double &Fx, double &Fy, double &Fz)
{
    if (KIND == POT_GENERIC) force_generic(P, ltab, x, y, z, Fx, Fy, Fz);
    else { force<KIND>(P, ltab, x, y, z, Fx, Fz); Fy = Fx; }
}

// t [Myr] is used by POT_GENERAL alone (time-interpolated amplitudes); it costs the other kinds nothing.
template <int KIND>
__device__ __forceinline__ void gradient(const PotParams &P, LogTab ltab, double t,
                                         double x, double y, double z,
                                         double &gx, double &gy, double &gz)
{
    if (KindTraits<KIND>::VEC) {
        gradient_vec<KIND>(P, ltab, t, x, y, z, gx, gy, gz);
    } else if (KindTraits<KIND>::TS) {
        double Fx, Fy, Fz;
        force3<KindTraits<KIND>::BASE>(P, ltab, x, y, z, Fx, Fy, Fz);
        const double f = amplitude_rel(P, 0, t);
        gx = (Fx * f) * x;
        gy = (Fy * f) * y;
        gz = (Fz * f) * z;
    } else {
        double Fx, Fy, Fz;
        force3<KIND>(P, ltab, x, y, z, Fx, Fy, Fz);
        gx = Fx * x;
        gy = Fy * y;
        gz = Fz * z;
    }
}

// The cubic of the knot interval a lane's clock is in, kept in registers by the leapfrog kernels of the separable
// kinds: an orbit crosses a knot once per Gyr or so, so the table look-up leaves the step.  Before the first and after
// the last knot the end values are held (constant segments reaching to -inf / +inf).
struct AmpSeg { double lo, hi, t0, c0, c1, c2, c3; };      // valid for lo <= t < hi

__device__ __forceinline__ void amp_seg_reset(AmpSeg &a) { a.lo = 1.0; a.hi = 0.0; a.t0 = a.c0 = a.c1 = a.c2 = a.c3 = 0.0; }

__device__ __forceinline__ void amp_seg_load(const PotParams &P, double t, AmpSeg &a)
{
    const int i0 = P.k_off[0], i1 = P.k_off[1];            // f(t) lives in the slot of "component 0"
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    a.t0 = 0.0; a.c1 = 0.0; a.c2 = 0.0; a.c3 = 0.0;
    if (i1 - i0 < 2 || !(t >= P.k_t[i0])) {
        a.lo = -INF; a.hi = (i1 - i0 < 2) ? INF : P.k_t[i0]; a.c0 = (i1 > i0) ? P.k_c[i0][0] : 1.0;
        return;
    }
    if (t >= P.k_t[i1 - 1]) {
        a.lo = P.k_t[i1 - 1]; a.hi = INF; a.c0 = P.k_c[i1 - 1][0];
        return;
    }
    int lo = i0, hi = i1 - 2;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t >= P.k_t[mid]) lo = mid; else hi = mid - 1;
    }
    a.lo = P.k_t[lo]; a.hi = P.k_t[lo + 1]; a.t0 = P.k_t[lo];
    a.c0 = P.k_c[lo][0]; a.c1 = P.k_c[lo][1]; a.c2 = P.k_c[lo][2]; a.c3 = P.k_c[lo][3];
}

__device__ __forceinline__ double amp_seg_eval(const PotParams &P, AmpSeg &a, double t)
{
    if (!(t >= a.lo && t < a.hi)) amp_seg_load(P, t, a);
    const double s = t - a.t0;
    return fma(s, fma(s, fma(s, a.c3, a.c2), a.c1), a.c0);
}

// The leapfrog's form.  Diagonal kinds: the three factors of grad = (Fx x, Fy y, Fz z), folded into the
// velocity FMAs by lf_kick.  POT_GENERAL: (Fx, Fy, Fz) IS the gradient.
template <int KIND>
__device__ __forceinline__ void force3t(const PotParams &P, LogTab ltab, double t,
                                        double x, double y, double z, double &Fx, double &Fy, double &Fz)
{
    if (KindTraits<KIND>::VEC) {
        gradient_vec<KIND>(P, ltab, t, x, y, z, Fx, Fy, Fz);
    } else if (KindTraits<KIND>::TS) {
        force3<KindTraits<KIND>::BASE>(P, ltab, x, y, z, Fx, Fy, Fz);
        const double f = amplitude_rel(P, 0, t);
        Fx *= f; Fy *= f; Fz *= f;
    } else {
        force3<KIND>(P, ltab, x, y, z, Fx, Fy, Fz);
    }
}

// force3t for the leapfrog kernels, which carry the current knot interval of a separable kind in registers
template <int KIND>
__device__ __forceinline__ void force3t_seg(const PotParams &P, LogTab ltab, AmpSeg &seg, double t,
                                            double x, double y, double z, double &Fx, double &Fy, double &Fz)
{
    if (KindTraits<KIND>::TS) {
        force3<KindTraits<KIND>::BASE>(P, ltab, x, y, z, Fx, Fy, Fz);
        const double f = amp_seg_eval(P, seg, t);
        Fx *= f; Fy *= f; Fz *= f;
    } else {
        force3t<KIND>(P, ltab, t, x, y, z, Fx, Fy, Fz);
    }
}

// vh + coef * grad_c, with grad_c = F * pos (diagonal kinds) or F itself (POT_GENERAL)
template <int KIND>
__device__ __forceinline__ double lf_kick(double F, double coef, double pos, double vh)
{
    return KindTraits<KIND>::VEC ? fma(F, coef, vh) : fma(F * coef, pos, vh);
}

// dt of a segment = t[1] - t[0] with both nodes converted to Myr first (gala converts the whole
// time array, then differences it).  The orbit is sensitive to the last bit of dt (a 2e-12
// relative change moves a 1 Gyr leapfrog orbit by ~1e-10), so the two products and the difference
// are rounded separately -- no FMA contraction.
__device__ __forceinline__ double seg_dt_myr(double t0, double t1, double to_myr)
{
    return __dsub_rn(__dmul_rn(t1, to_myr), __dmul_rn(t0, to_myr));
}

// Phi of the same components (gala *_value functions) -- for the N3 rows, not the hot loop.
__device__ __forceinline__ double potential_value(const PotParams &P, double x0, double y0, double z0, double t_eval)
{
    double vtot = 0.0;
    for (int i = 0; i < P.n_comp; i++) {
        const int t = P.type[i];
        double x = x0, y = y0, z = z0, rel = 1.0;
        if (P.kind == POT_MW_V2_T || P.kind == POT_GENERIC_T) rel = amplitude_rel(P, 0, t_eval);
        if (P.kind == POT_GENERAL || P.kind == POT_ROTATED || P.kind == POT_LM10) {
            rel = amplitude_rel(P, i, t_eval);
            if (P.has_R[i]) {
                x = fma(P.R[i][2], z0, fma(P.R[i][1], y0, P.R[i][0] * x0));
                y = fma(P.R[i][5], z0, fma(P.R[i][4], y0, P.R[i][3] * x0));
                z = fma(P.R[i][8], z0, fma(P.R[i][7], y0, P.R[i][6] * x0));
            }
        }
        const double R2 = fma(y, y, x * x);
        const double z2 = z * z;
        const double r = sqrt(R2 + z2);
        double v = 0.0;
        if (comp_is_nondiag(t)) {
            double cx, cy, cz;
            nondiag_component(P, i, P.GM[i], x, y, z, cx, cy, cz, &v);
        } else if (t == CW_COMP_MIYAMOTO_NAGAI) {
            const double zd = P.p1[i] + sqrt(fma(P.p2[i], P.p2[i], z2));
            v -= P.GM[i] / sqrt(fma(zd, zd, R2));
        } else if (t == CW_COMP_HERNQUIST) {
            v -= P.GM[i] / (r + P.p1[i]);
        } else if (t == CW_COMP_LOG_TRIAXIAL) {
            v += 0.5 * P.GM[i] * log(fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], fma(x * x, P.q2[i][0], P.p1[i]))));
        } else if (t == CW_COMP_NFW_TRIAXIAL) {
            const double u = sqrt(fma(z2, P.q2[i][2], fma(y * y, P.q2[i][1], (x * x) * P.q2[i][0]))) * P.p2[i];
            v -= (P.GM[i] * P.p2[i]) * log(1.0 + u) / u;
        } else if (t == CW_COMP_KUZMIN) {
            const double az = P.p1[i] + fabs(z);
            v -= P.GM[i] / sqrt(fma(az, az, R2));
        } else if (t == CW_COMP_BURKERT) {
            const double xb = r * P.p2[i], ix = 1.0 / xb;
            v -= P.GM[i] * P.p1[i] * (3.14159265358979323846 - 2.0 * (1.0 + ix) * atan(xb) + 2.0 * (1.0 + ix) * log(1.0 + xb)
                                      - (1.0 - ix) * log(fma(xb, xb, 1.0)));
        } else if (t == CW_COMP_KEPLER) {
            v -= P.GM[i] / r;
        } else if (t == CW_COMP_PLUMMER) {
            v -= P.GM[i] / sqrt(R2 + z2 + P.p2[i]);
        } else if (t == CW_COMP_ISOCHRONE) {
            v -= P.GM[i] / (sqrt(R2 + z2 + P.p2[i]) + P.p1[i]);
        } else if (t == CW_COMP_JAFFE) {
            v += P.GM[i] / P.p1[i] * log(r / (r + P.p1[i]));
        } else if (t == CW_COMP_LOGARITHMIC) {
            v += 0.5 * P.GM[i] * log(fma(z2, P.p2[i], R2 + P.p1[i]));
        } else {
            const double u = r / P.p1[i];
            v -= (P.GM[i] / P.p1[i]) * log(1.0 + u) / u;
        }
        vtot = fma(v, rel, vtot);
    }
    return vtot;
}

} // namespace cw
