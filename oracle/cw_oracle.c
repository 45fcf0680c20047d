/*
 * cw_oracle.c -- CPU ORACLE for the cogsworth galactic-dynamics hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (cogsworth_b200/) never links, imports or calls anything in oracle/.
 *
 * It restates, in plain C with gala's operation order, the algorithm that
 *   cogsworth.kicks.integrate_orbit_with_events   (/root/reference/src/cogsworth/kicks.py:52-209)
 * runs per orbit on the CPU:  time grid -> segment-by-segment gala integration ->
 * velocity kick at each supernova -> stitch.
 *
 * The arithmetic itself lives in a third-party dependency that is NOT under /root/reference
 * and is not installed here:  gala (PyPI `gala`, pinned only as `>= 1.11.0`,
 * /root/reference/pyproject.toml:24).  The functions below therefore restate gala's
 * published algorithms (file names as in the gala source tree):
 *   gala/integrate/timespec.py                  parse_time_specification   -> cwo_time_grid
 *   gala/potential/potential/builtin/builtin_potentials.c
 *        hernquist_gradient / miyamotonagai_gradient / mn3_gradient / sphericalnfw_gradient
 *                                                                        -> cwo_gradient
 *   gala/potential/potential/src/cpotential.c   c_gradient (component loop) -> cwo_gradient
 *   gala/integrate/cyintegrators/leapfrog.pyx   c_init_velocity / c_leapfrog_step -> cwo_leapfrog
 *   gala/integrate/cyintegrators/dop853.pyx     dop853_integrate_hamiltonian  -> cwo_dop853_segment
 *   gala/integrate/cyintegrators/dopri/dop853.c Hairer's dopcor            -> cwo_dop853
 *
 * PARITY PINNING: the reference's own tests hold no golden trajectories.  The
 * oracle is pinned against the numeric outputs printed in the reference's executed notebook
 * docs/tutorials/visualisation/gala.ipynb (circular velocity, time grid, six phase-space
 * samples of a DOPRI853 orbit), the tests/test_events.py fixture, and scipy's compiled Hairer
 * DOP853 (tests/test_oracle_*.py).  It has NOT been compared with gala run here, because gala
 * cannot be installed in this container: at the gala boundary parity is "unpinned".
 *
 * Build:  gcc -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off keeps gcc from fusing a*b+c, matching gala's x86-64 wheels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/cw_dop853_coeffs.h"

#define CWO_MAX_COMP 16
#define CWO_COMP_MN 1
#define CWO_COMP_HERNQUIST 2
#define CWO_COMP_NFW 3
#define CWO_COMP_KEPLER 4       /* (m, -, -) */
#define CWO_COMP_PLUMMER 5      /* (m, b, -) */
#define CWO_COMP_ISOCHRONE 6    /* (m, b, -) */
#define CWO_COMP_JAFFE 7        /* (m, c, -) */
#define CWO_COMP_LOGARITHMIC 8  /* (v_c, r_h, q3) with q1 = q2 = 1 (axisymmetric) */
#define CWO_COMP_KUZMIN 9       /* (m, a, -) */
#define CWO_COMP_BURKERT 10     /* (rho, r0, -) */
#define CWO_COMP_LOG_TRIAXIAL 12 /* (v_c, r_h, q1, q2, q3), phi = 0: LogarithmicPotential with axis ratios */
#define CWO_COMP_NFW_TRIAXIAL 11 /* (m, r_s, a, b, c): NFWPotential with flattening (docs/tutorials/pop_settings/potentials.ipynb:141) */

#define CWO_LEAPFROG 0
#define CWO_DOP853 1        /* restarted on every grid interval */
#define CWO_DOP853_DENSE 2  /* one call per segment, dense output at the grid nodes */

#define CWO_FLAG_APPEND_T2 1

/* per-orbit status, Hairer's return codes plus two of ours */
#define CWO_OK 0
#define CWO_E_INCONSISTENT (-1)
#define CWO_E_NMAX (-2)
#define CWO_E_STEP_TOO_SMALL (-3)
#define CWO_E_STIFF (-4)
#define CWO_E_NONFINITE (-5)
#define CWO_E_CAPACITY (-6)

#define CWO_COMP_LONG_MURALI_BAR 13 /* (m, a, b, c); the angle alpha is a rotation */
#define CWO_COMP_LEESUTO_NFW 14     /* (v_c, r_s, a, b, c) */
#define CWO_COMP_POWERLAW_CUTOFF 15 /* (m, alpha, r_c) */
#define CWO_MAX_KNOTS 32            /* per component */

typedef struct {
    double G;
    int32_t n;
    int32_t type[CWO_MAX_COMP];
    double p[CWO_MAX_COMP][6];
    /* gala's generic per-potential rotation (cpotential.c applies R to q, R^T to the gradient) */
    int32_t has_R[CWO_MAX_COMP];
    double R[CWO_MAX_COMP][9];
    /* gp.TimeInterpolatedPotential: p[i][0] at the knots; interp 0 = linear, 1 = natural cubic spline (GSL cspline) */
    int32_t n_knots[CWO_MAX_COMP];
    int32_t interp;
    double knot_t[CWO_MAX_COMP][CWO_MAX_KNOTS];
    double knot_y[CWO_MAX_COMP][CWO_MAX_KNOTS];
    double knot_c[CWO_MAX_COMP][CWO_MAX_KNOTS];   /* GSL cspline's c_i (half second derivatives), by cwo_prepare */
    double t_eval;                                 /* time of cwo_gradient / cwo_value (the integrators pass their own) */
} cwo_potential;

typedef struct {
    int64_t n_accepted, n_rejected, n_fcn;
} cwo_stats;

/* ------------------------------------------------------------------ potentials */

/* builtin_potentials.c: miyamotonagai_gradient, pars = (G, m, a, b) */
static void mn_gradient(double G, const double *p, const double *q, double *grad)
{
    double sqrtz = sqrt(q[2] * q[2] + p[2] * p[2]);
    double zd = p[1] + sqrtz;
    double fac = G * p[0] * pow(q[0] * q[0] + q[1] * q[1] + zd * zd, -1.5);
    grad[0] = grad[0] + fac * q[0];
    grad[1] = grad[1] + fac * q[1];
    grad[2] = grad[2] + fac * q[2] * (1. + p[1] / sqrtz);
}

/* builtin_potentials.c: hernquist_gradient, pars = (G, m, c) */
static void hernquist_gradient(double G, const double *p, const double *q, double *grad)
{
    double R = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    double fac = G * p[0] / ((R + p[1]) * (R + p[1]) * R);
    grad[0] = grad[0] + fac * q[0];
    grad[1] = grad[1] + fac * q[1];
    grad[2] = grad[2] + fac * q[2];
}

/* builtin_potentials.c: sphericalnfw_gradient, pars = (G, m, r_s) */
static void nfw_gradient(double G, const double *p, const double *q, double *grad)
{
    double v_h2 = G * p[0] / p[1];
    double u = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) / p[1];
    double fac = v_h2 / (u * u * u) / (p[1] * p[1]) * (log(1 + u) - u / (1 + u));
    grad[0] = grad[0] + fac * q[0];
    grad[1] = grad[1] + fac * q[1];
    grad[2] = grad[2] + fac * q[2];
}

/* The remaining spherical / axisymmetric builtins a cogsworth user can put into a CompositePotential
 * (online-interface/potentials.py:93-131), restated from builtin_potentials.c [gala-recall]:
 *   kepler_gradient      fac = G m / r^3
 *   plummer_gradient     fac = G m / (r^2 + b^2)^1.5
 *   isochrone_gradient   fac = G m / (s (s + b)^2),  s = sqrt(r^2 + b^2)
 *   jaffe_gradient       fac = G m / (r^2 (r + c))          [Phi = G m / c ln(r / (r + c))]
 *   logarithmic_gradient fac = v_c^2 / (r_h^2 + x^2/q1^2 + y^2/q2^2 + z^2/q3^2), grad_i = fac q_i / q_i^2
 *                        (only q1 = q2 = 1: the kernels factor the gradient as (Fxy x, Fxy y, Fz z))
 *   kuzmin_gradient      fac = G m / (R^2 + (a + |z|)^2)^1.5, grad = (fac x, fac y, fac (a + |z|) sign z)
 *   burkert_gradient     dPhi/dr = pi G rho r0 (ln(1 + x^2) + 2 ln(1 + x) - 2 atan x) / x^2,  x = r / r0 */
static void extra_gradient(int type, double G, const double *p, const double *q, double *grad)
{
    double r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    double fxy, fz;
    if (type == CWO_COMP_KEPLER) {
        double R = sqrt(r2);
        fxy = fz = G * p[0] / (R * R * R);
    } else if (type == CWO_COMP_PLUMMER) {
        fxy = fz = G * p[0] / pow(r2 + p[1] * p[1], 1.5);
    } else if (type == CWO_COMP_ISOCHRONE) {
        double s = sqrt(r2 + p[1] * p[1]);
        fxy = fz = G * p[0] / (s * (s + p[1]) * (s + p[1]));
    } else if (type == CWO_COMP_JAFFE) {
        double R = sqrt(r2);
        fxy = fz = G * p[0] / (R * R * (R + p[1]));
    } else if (type == CWO_COMP_KUZMIN) {
        double az = p[1] + fabs(q[2]);
        double fac = G * p[0] / pow(q[0] * q[0] + q[1] * q[1] + az * az, 1.5);
        fxy = fac;
        fz = (q[2] != 0.0) ? fac * az / fabs(q[2]) : 0.0;    /* fz z = fac (a + |z|) sign z */
    } else if (type == CWO_COMP_BURKERT) {
        double R = sqrt(r2), x = R / p[1];
        double dphi_dr = M_PI * G * p[0] * p[1] / (x * x) * (log(1. + x * x) + 2. * log(1. + x) - 2. * atan(x));
        fxy = fz = dphi_dr / R;
    } else { /* CWO_COMP_LOGARITHMIC */
        double fac = p[0] * p[0] / (p[1] * p[1] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2] / (p[2] * p[2]));
        fxy = fac;
        fz = fac / (p[2] * p[2]);
    }
    grad[0] = grad[0] + fxy * q[0];
    grad[1] = grad[1] + fxy * q[1];
    grad[2] = grad[2] + fz * q[2];
}

/* builtin_potentials.c: triaxialnfw_gradient, pars = (G, m, r_s, a, b, c) [gala-recall]:
 * the spherical formula in the flattened radius u = sqrt(x^2/a^2 + y^2/b^2 + z^2/c^2) / r_s,
 * gradient components divided by a^2, b^2, c^2 */
static void triaxial_nfw_gradient(double G, const double *p, const double *q, double *grad)
{
    double v_h2 = G * p[0] / p[1];
    double u = sqrt(q[0] * q[0] / (p[2] * p[2]) + q[1] * q[1] / (p[3] * p[3]) + q[2] * q[2] / (p[4] * p[4])) / p[1];
    double fac = v_h2 / (u * u * u) / (p[1] * p[1]) * (log(1 + u) - u / (1 + u));
    grad[0] = grad[0] + fac * q[0] / (p[2] * p[2]);
    grad[1] = grad[1] + fac * q[1] / (p[3] * p[3]);
    grad[2] = grad[2] + fac * q[2] / (p[4] * p[4]);
}

/* logarithmic_gradient with q1, q2, q3 and no rotation (phi = 0) [gala-recall]:
 * fac = v_c^2 / (r_h^2 + x^2/q1^2 + y^2/q2^2 + z^2/q3^2), grad_i = fac q_i / q_i^2 */
static void triaxial_log_gradient(const double *p, const double *q, double *grad)
{
    double fac = p[0] * p[0] / (p[1] * p[1] + q[0] * q[0] / (p[2] * p[2]) + q[1] * q[1] / (p[3] * p[3]) + q[2] * q[2] / (p[4] * p[4]));
    grad[0] = grad[0] + fac * q[0] / (p[2] * p[2]);
    grad[1] = grad[1] + fac * q[1] / (p[3] * p[3]);
    grad[2] = grad[2] + fac * q[2] / (p[4] * p[4]);
}

/* builtin_potentials.c: longmuralibar_gradient, pars = (G, m, a, b, c) in the bar's own frame [gala-recall]
 * (Long & Murali 1992: Phi = G m / (2a) ln((x - a + T-) / (x + a + T+)), T+- = sqrt((a +- x)^2 + y^2 + (b + sqrt(c^2 + z^2))^2)) */
static double longmurali_value(double G, const double *p, const double *q)
{
    double a = p[1], b = p[2], c = p[3];
    double bz = b + sqrt(c * c + q[2] * q[2]);
    double Tm = sqrt((a - q[0]) * (a - q[0]) + q[1] * q[1] + bz * bz);
    double Tp = sqrt((a + q[0]) * (a + q[0]) + q[1] * q[1] + bz * bz);
    return G * p[0] / (2. * a) * log((q[0] - a + Tm) / (q[0] + a + Tp));
}
static void longmurali_gradient(double G, const double *p, const double *q, double *grad)
{
    double a = p[1], b = p[2], c = p[3];
    double x = q[0], y = q[1], z = q[2];
    double sz = sqrt(c * c + z * z), bz = b + sz;
    double Tm = sqrt((a - x) * (a - x) + y * y + bz * bz);
    double Tp = sqrt((a + x) * (a + x) + y * y + bz * bz);
    double fac1 = G * p[0] / (2. * Tm * Tp);
    double fac2 = 1. / (y * y + bz * bz);
    double fac3 = Tm + Tp - 4. * x * x / (Tm + Tp);
    grad[0] = grad[0] + 4. * fac1 * x / (Tm + Tp);
    grad[1] = grad[1] + fac1 * y * fac2 * fac3;
    grad[2] = grad[2] + fac1 * z * fac2 * fac3 * bz / sz;
}

/* builtin_potentials.c: leesuto_*, pars = (G, v_c, r_s, a, b, c) [gala-recall] (Lee & Suto 2003, eq. 27-29):
 * Phi = phi0 (F1(u) + (e_b^2 + e_c^2)/2 F2(u) + (e_b^2 y^2 + e_c^2 z^2) / (2 r^2) F3(u)), u = r / r_s */
static double leesuto_value(const double *p, const double *q)
{
    double eb2 = 1. - (p[3] / p[2]) * (p[3] / p[2]), ec2 = 1. - (p[4] / p[2]) * (p[4] / p[2]);
    double phi0 = p[0] * p[0] / (log(2.) - 0.5 + (log(2.) - 0.75) * eb2 + (log(2.) - 0.75) * ec2);
    double r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], r = sqrt(r2), u = r / p[1];
    double F1 = -log(1. + u) / u;
    double F2 = -1. / 3. + (2. * u * u - 3. * u + 6.) / (6. * u * u) + (1. / u - 1. / (u * u * u)) * log(1. + u);
    double F3 = (u * u - 3. * u - 6.) / (2. * u * u * (1. + u)) + 3. / (u * u * u) * log(1. + u);
    return phi0 * (F1 + (eb2 + ec2) / 2. * F2 + (eb2 * q[1] * q[1] + ec2 * q[2] * q[2]) / (2. * r2) * F3);
}
static void leesuto_gradient(const double *p, const double *q, double *grad)
{
    double eb2 = 1. - (p[3] / p[2]) * (p[3] / p[2]), ec2 = 1. - (p[4] / p[2]) * (p[4] / p[2]);
    double phi0 = p[0] * p[0] / (log(2.) - 0.5 + (log(2.) - 0.75) * eb2 + (log(2.) - 0.75) * ec2);
    double x = q[0], y = q[1], z = q[2];
    double r2 = x * x + y * y + z * z, r = sqrt(r2), u = r / p[1], L = log(1. + u);
    double u2 = u * u, u3 = u2 * u, u4 = u2 * u2;
    double F3 = (u2 - 3. * u - 6.) / (2. * u2 * (1. + u)) + 3. * L / u3;
    double dF1 = L / u2 - 1. / (u * (1. + u));
    double dF2 = 1. / (2. * u2) - 2. / u3 + (3. / u4 - 1. / u2) * L + (1. / u - 1. / u3) / (1. + u);
    double num = u2 - 3. * u - 6., den = 2. * u2 * (1. + u);
    double dF3 = ((2. * u - 3.) * den - num * (4. * u + 6. * u2)) / (den * den) + 3. / (u3 * (1. + u)) - 9. * L / u4;
    double S = (eb2 * y * y + ec2 * z * z) / 2.;
    double dPhi_du = dF1 + (eb2 + ec2) / 2. * dF2 + S / r2 * dF3;      /* at fixed angles */
    double rad = dPhi_du / (p[1] * r);                                /* du/dq_i = q_i / (r r_s) */
    grad[0] = grad[0] + phi0 * (rad * x + F3 * (-2. * S * x / (r2 * r2)));
    grad[1] = grad[1] + phi0 * (rad * y + F3 * (eb2 * y / r2 - 2. * S * y / (r2 * r2)));
    grad[2] = grad[2] + phi0 * (rad * z + F3 * (ec2 * z / r2 - 2. * S * z / (r2 * r2)));
}

/* regularised lower incomplete gamma P(a, x) (what gala takes from GSL: gsl_sf_gamma_inc_P): series / continued fraction */
static double gamma_inc_P(double a, double x)
{
    if (!(x > 0.)) return 0.;
    double pref = exp(a * log(x) - x - lgamma(a));
    if (x < a + 1.) {
        double ap = a, del = 1. / a, sum = del;
        for (int n = 0; n < 1000; n++) { ap += 1.; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
        return sum * pref;
    }
    double tiny = 1e-300, b = x + 1. - a, c = 1. / tiny, d = 1. / b, h = d;
    for (int n = 1; n < 1000; n++) {
        double an = -(double)n * ((double)n - a);
        b += 2.;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1. / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.) < 1e-16) break;
    }
    return 1. - pref * h;
}

/* builtin_potentials.c: powerlawcutoff_gradient, pars = (G, m, alpha, r_c) [gala-recall]: rho ~ r^-alpha exp(-r^2/r_c^2)
 * of total mass m, dPhi/dr = G M(<r) / r^2 with M(<r) = m P(3/2 - alpha/2, r^2 / r_c^2) */
static void powerlawcutoff_gradient(double G, const double *p, const double *q, double *grad)
{
    double r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], r = sqrt(r2);
    double dPhi_dr = G * p[0] * gamma_inc_P(1.5 - p[1] / 2., r2 / (p[2] * p[2])) / r2;
    grad[0] = grad[0] + dPhi_dr * q[0] / r;
    grad[1] = grad[1] + dPhi_dr * q[1] / r;
    grad[2] = grad[2] + dPhi_dr * q[2] / r;
}
/* Phi = -G M(<r) / r - G m Gamma(1 - alpha/2, r^2/r_c^2) / (r_c Gamma(3/2 - alpha/2)) (zero at infinity; alpha < 2) */
static double powerlawcutoff_value(double G, const double *p, const double *q)
{
    double r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], r = sqrt(r2);
    double s = 1.5 - p[1] / 2., x = r2 / (p[2] * p[2]);
    if (!(s > 0.5)) return NAN;
    double Q2 = 1. - gamma_inc_P(s - 0.5, x);
    return -G * p[0] * (gamma_inc_P(s, x) / r + Q2 * exp(lgamma(s - 0.5) - lgamma(s)) / p[2]);
}

/* Natural cubic spline exactly as GSL's cspline.c sets it up: c_0 = c_{n-1} = 0, the interior c_i from the
 * tridiagonal system; evaluation y = y_i + dx (b_i + dx (c_i + dx d_i)). */
void cwo_prepare(cwo_potential *P)
{
    for (int k = 0; k < P->n; k++) {
        int n = P->n_knots[k];
        double *c = P->knot_c[k];
        const double *x = P->knot_t[k], *y = P->knot_y[k];
        for (int i = 0; i < CWO_MAX_KNOTS; i++) c[i] = 0.;
        if (n < 3 || P->interp == 0) continue;
        int m = n - 2;                       /* unknowns c_1 .. c_{n-2} */
        double diag[CWO_MAX_KNOTS], off[CWO_MAX_KNOTS], g[CWO_MAX_KNOTS];
        for (int i = 0; i < m; i++) {
            double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
            double yd_i = y[i + 1] - y[i], yd_ip1 = y[i + 2] - y[i + 1];
            off[i] = h_ip1;
            diag[i] = 2. * (h_ip1 + h_i);
            g[i] = 3. * (yd_ip1 / h_ip1 - yd_i / h_i);
        }
        /* symmetric tridiagonal solve (Gaussian elimination without pivoting: diagonally dominant) */
        for (int i = 1; i < m; i++) {
            double w = off[i - 1] / diag[i - 1];
            diag[i] -= w * off[i - 1];
            g[i] -= w * g[i - 1];
        }
        for (int i = m - 1; i >= 0; i--) {
            double v = g[i];
            if (i + 1 < m) v -= off[i] * c[i + 2];
            c[i + 1] = v / diag[i];
        }
    }
}

/* the first parameter of component k at time t; outside the knots the end values are held */
static double first_parameter(const cwo_potential *P, int k, double t)
{
    int n = P->n_knots[k];
    if (n <= 0) return P->p[k][0];
    const double *x = P->knot_t[k], *y = P->knot_y[k], *c = P->knot_c[k];
    if (n == 1 || t <= x[0]) return y[0];
    if (t >= x[n - 1]) return y[n - 1];
    int i = 0;
    while (i + 2 < n && t >= x[i + 1]) i++;
    double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i], d = t - x[i];
    if (P->interp == 0) return y[i] + d / dx * dy;
    double b_i = dy / dx - dx * (c[i + 1] + 2. * c[i]) / 3.;
    double d_i = (c[i + 1] - c[i]) / (3. * dx);
    return y[i] + d * (b_i + d * (c[i] + d * d_i));
}

static void component_gradient(int type, double G, const double *p, const double *q, double *grad)
{
    switch (type) {
    case CWO_COMP_MN: mn_gradient(G, p, q, grad); break;
    case CWO_COMP_HERNQUIST: hernquist_gradient(G, p, q, grad); break;
    case CWO_COMP_NFW: nfw_gradient(G, p, q, grad); break;
    case CWO_COMP_NFW_TRIAXIAL: triaxial_nfw_gradient(G, p, q, grad); break;
    case CWO_COMP_LOG_TRIAXIAL: triaxial_log_gradient(p, q, grad); break;
    case CWO_COMP_KEPLER: case CWO_COMP_PLUMMER: case CWO_COMP_ISOCHRONE: case CWO_COMP_JAFFE:
    case CWO_COMP_LOGARITHMIC: case CWO_COMP_KUZMIN: case CWO_COMP_BURKERT:
        extra_gradient(type, G, p, q, grad); break;
    case CWO_COMP_LONG_MURALI_BAR: longmurali_gradient(G, p, q, grad); break;
    case CWO_COMP_LEESUTO_NFW: leesuto_gradient(p, q, grad); break;
    case CWO_COMP_POWERLAW_CUTOFF: powerlawcutoff_gradient(G, p, q, grad); break;
    default: grad[0] = grad[1] = grad[2] = NAN;
    }
}

/* Rounding-noise mode (tests only; off by default).  With eps != 0 every gradient component is multiplied by
 * 1 + eps * s, s in {-1, 0, +1} drawn from a hash of the position bits and the seed: the arithmetic of ANY equivalent
 * but differently rounded implementation (another libm, FMA contraction, a GPU's rsqrt / log) differs from this file by
 * such last-bit factors at every evaluation.  Two runs that differ only in this noise tell which orbits amplify
 * last-bit differences beyond the parity tolerance -- no implementation can agree with another on those. */
static double g_noise_eps = 0.0;
static uint64_t g_noise_seed = 0;
void cwo_set_rounding_noise(double eps, uint64_t seed) { g_noise_eps = eps; g_noise_seed = seed; }

static void apply_rounding_noise(const double *q, double *grad)
{
    uint64_t h = g_noise_seed * 0x9E3779B97F4A7C15ull;
    for (int a = 0; a < 3; a++) {
        uint64_t b;
        memcpy(&b, &q[a], 8);
        h ^= b + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    for (int a = 0; a < 3; a++) {
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        const int s = (int)(h % 3) - 1;
        grad[a] = grad[a] * (1.0 + g_noise_eps * (double)s);
    }
}

static void gradient_exact(const cwo_potential *P, double t, const double *q, double *grad);

/* cpotential.c: c_gradient -- zero, then accumulate components in order; a component with a rotation sees
 * q' = R q and contributes R^T grad'; a time-interpolated one is evaluated with its parameters at time t */
void cwo_gradient_t(const cwo_potential *P, double t, const double *q, double *grad)
{
    gradient_exact(P, t, q, grad);
    if (g_noise_eps != 0.0) apply_rounding_noise(q, grad);
}

static void gradient_exact(const cwo_potential *P, double t, const double *q, double *grad)
{
    grad[0] = 0.; grad[1] = 0.; grad[2] = 0.;
    for (int i = 0; i < P->n; i++) {
        double p[6];
        for (int k = 0; k < 6; k++) p[k] = P->p[i][k];
        p[0] = first_parameter(P, i, t);
        if (P->has_R[i]) {
            const double *R = P->R[i];
            double qr[3], gr[3] = {0., 0., 0.};
            for (int a = 0; a < 3; a++) qr[a] = R[3 * a] * q[0] + R[3 * a + 1] * q[1] + R[3 * a + 2] * q[2];
            component_gradient(P->type[i], P->G, p, qr, gr);
            for (int a = 0; a < 3; a++) grad[a] = grad[a] + (R[a] * gr[0] + R[3 + a] * gr[1] + R[6 + a] * gr[2]);
        } else {
            component_gradient(P->type[i], P->G, p, q, grad);
        }
    }
}

void cwo_gradient(const cwo_potential *P, const double *q, double *grad)
{
    cwo_gradient_t(P, P->t_eval, q, grad);
}

/* builtin_potentials.c *_value: Phi of the same components (used by the "next" rows N3) */
double cwo_value(const cwo_potential *P, const double *q0)
{
    double v = 0.;
    for (int i = 0; i < P->n; i++) {
        double p[6], q[3] = {q0[0], q0[1], q0[2]};
        for (int k = 0; k < 6; k++) p[k] = P->p[i][k];
        p[0] = first_parameter(P, i, P->t_eval);
        if (P->has_R[i]) {
            const double *R = P->R[i];
            for (int a = 0; a < 3; a++) q[a] = R[3 * a] * q0[0] + R[3 * a + 1] * q0[1] + R[3 * a + 2] * q0[2];
        }
        if (P->type[i] == CWO_COMP_LONG_MURALI_BAR) v += longmurali_value(P->G, p, q);
        else if (P->type[i] == CWO_COMP_LEESUTO_NFW) v += leesuto_value(p, q);
        else if (P->type[i] == CWO_COMP_POWERLAW_CUTOFF) v += powerlawcutoff_value(P->G, p, q);
        else if (P->type[i] == CWO_COMP_MN) {
            double zd = p[1] + sqrt(q[2] * q[2] + p[2] * p[2]);
            v += -P->G * p[0] / sqrt(q[0] * q[0] + q[1] * q[1] + zd * zd);
        } else if (P->type[i] == CWO_COMP_HERNQUIST) {
            double R = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            v += -P->G * p[0] / (R + p[1]);
        } else if (P->type[i] == CWO_COMP_NFW) {
            double u = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) / p[1];
            double v_h2 = -P->G * p[0] / p[1];
            v += v_h2 * log(1 + u) / u;
        } else if (P->type[i] == CWO_COMP_NFW_TRIAXIAL) {
            double u = sqrt(q[0] * q[0] / (p[2] * p[2]) + q[1] * q[1] / (p[3] * p[3]) + q[2] * q[2] / (p[4] * p[4])) / p[1];
            v += -P->G * p[0] / p[1] * log(1 + u) / u;
        } else if (P->type[i] == CWO_COMP_LOG_TRIAXIAL) {
            v += 0.5 * p[0] * p[0] * log(p[1] * p[1] + q[0] * q[0] / (p[2] * p[2]) + q[1] * q[1] / (p[3] * p[3]) + q[2] * q[2] / (p[4] * p[4]));
        } else if (P->type[i] == CWO_COMP_KUZMIN) {
            double az = p[1] + fabs(q[2]);
            v += -P->G * p[0] / sqrt(q[0] * q[0] + q[1] * q[1] + az * az);
        } else if (P->type[i] == CWO_COMP_BURKERT) {
            /* Mori & Burkert 2000, as gala's burkert_value */
            double x = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) / p[1];
            v += -M_PI * P->G * p[0] * p[1] * p[1] * (M_PI - 2. * (1. + 1. / x) * atan(x) + 2. * (1. + 1. / x) * log(1. + x)
                                                     - (1. - 1. / x) * log(1. + x * x));
        } else if (P->type[i] >= CWO_COMP_KEPLER && P->type[i] <= CWO_COMP_LOGARITHMIC) {
            double r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], R = sqrt(r2);
            switch (P->type[i]) {
            case CWO_COMP_KEPLER: v += -P->G * p[0] / R; break;
            case CWO_COMP_PLUMMER: v += -P->G * p[0] / sqrt(r2 + p[1] * p[1]); break;
            case CWO_COMP_ISOCHRONE: v += -P->G * p[0] / (sqrt(r2 + p[1] * p[1]) + p[1]); break;
            case CWO_COMP_JAFFE: v += P->G * p[0] / p[1] * log(R / (R + p[1])); break;
            default: v += 0.5 * p[0] * p[0] * log(p[1] * p[1] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2] / (p[2] * p[2]));
            }
        } else v = NAN;
    }
    return v;
}

/* PotentialBase.circular_velocity: sqrt(r * |dPhi/dr|), dPhi/dr = sum(q * grad) / r */
double cwo_circular_velocity(const cwo_potential *P, const double *q)
{
    double g[3];
    cwo_gradient(P, q, g);
    double r = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    double dPhi_dr = (q[0] * g[0] + q[1] * g[1] + q[2] * g[2]) / r;
    return sqrt(r * fabs(dPhi_dr));
}

/* ------------------------------------------------------------------ time grid */

/* timespec.py parse_time_specification(dt, t1, t2), forward branch:
 *     t_i = t1; while (t_i < t2) and (ii < 1e6): times.append(t_i); t_i += dt; ii += 1
 * followed by kicks.py:121-122 (append t2 if the last node is < t2) when `append_t2`.
 * Returns the number of nodes; writes at most `cap` of them to T (T may be NULL to count). */
int64_t cwo_time_grid(double t1, double t2, double dt, int append_t2, double *T, int64_t cap)
{
    int64_t n = 0;
    double t_i = t1, last = t1;
    if (dt < 0.0 && t2 < t1) {
        /* backward branch of parse_time_specification [gala-recall]:
         *     while (t_i > t2) and (ii < 1e6): times.append(t_i); t_i += dt
         *     if times[-1] != t2: times.append(t2)
         * gala appends t2 itself here (it does not on the forward branch); the caller asks for it with
         * `append_t2`, as cogsworth.hydro.rewind._tohspans implies (hydro/rewind.py:16-19). */
        while (t_i > t2 && n < 1000000) {
            if (T && n < cap) T[n] = t_i;
            last = t_i;
            n++;
            t_i += dt;
        }
        if (n == 0) return 0;
        if (append_t2 && last != t2) {
            if (T && n < cap) T[n] = t2;
            n++;
        }
        return n;
    }
    if (!(dt > 0.0)) return 0;      /* dt <= 0 with t2 >= t1: gala raises ValueError */
    while (t_i < t2 && n < 1000000) {
        if (T && n < cap) T[n] = t_i;
        last = t_i;
        n++;
        t_i += dt;
    }
    if (n == 0) return 0;           /* t2 <= t1: gala raises ValueError */
    if (append_t2 && last < t2) {
        if (T && n < cap) T[n] = t2;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ leapfrog */

/* leapfrog.pyx leapfrog_integrate_hamiltonian for ONE orbit over t[0..nt-1] (Myr).
 * w = (x,y,z,vx,vy,vz) is advanced in place to t[nt-1]; when `out` != NULL sample j is
 * written to out[c*stride + j] for c = 0..5 (sample 0 is the input state).
 * nt == 1 is returned unchanged (gala reads t[1] out of bounds but never steps). */
void cwo_leapfrog(const cwo_potential *P, double *w, const double *t, int64_t nt,
                  double *out, int64_t stride)
{
    double grad[3], vh[3];
    if (out) for (int c = 0; c < 6; c++) out[c * stride] = w[c];
    if (nt < 2) return;
    double dt = t[1] - t[0];
    /* c_init_velocity at t[0]; every c_leapfrog_step evaluates the gradient at the new node's time t[j] [gala-recall] */
    cwo_gradient_t(P, t[0], w, grad);
    for (int k = 0; k < 3; k++) vh[k] = w[3 + k] - grad[k] * dt / 2.;
    for (int64_t j = 1; j < nt; j++) {
        /* c_leapfrog_step */
        for (int k = 0; k < 3; k++) w[k] = w[k] + vh[k] * dt;
        cwo_gradient_t(P, t[j], w, grad);
        for (int k = 0; k < 3; k++) {
            w[3 + k] = vh[k] - grad[k] * dt / 2.;
            vh[k] = vh[k] - grad[k] * dt;
        }
        if (out) for (int c = 0; c < 6; c++) out[c * stride + j] = w[c];
    }
}

/* ------------------------------------------------------------------ DOP853 (Hairer) */

#define NEQ 6

/* Fwrapper in gala's dop853.pyx: f = (dH/dp, -dH/dq) = (v, -grad Phi) */
static void fcn(const cwo_potential *P, double t, const double *y, double *f)
{
    double g[3];
    cwo_gradient_t(P, t, y, g);
    f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
    f[3] = -g[0]; f[4] = -g[1]; f[5] = -g[2];
}

static double sign_d(double a, double b) { return (b > 0.0) ? fabs(a) : -fabs(a); }
static double min_d(double a, double b) { return (a < b) ? a : b; }
static double max_d(double a, double b) { return (a > b) ? a : b; }

/* Hairer dop853.c: dop853() argument checks + dopcor(), specialised to what gala passes:
 * itoler=0 (scalar tolerances), iout=0, uround/safe/fac1/fac2/beta/hmax = 0 -> defaults,
 * h = dt0 (non-zero so hinit is skipped), meth=0 -> 1, nstiff=1, no dense output.
 * Integrates y from x to xend in place.  Returns 1 on success or a negative Hairer code. */
/* Dense output (iout = 2, contd8): when nout > 0, the solution at every requested time
 * tout[j] (ascending in the direction of integration, all inside (x, xend)) is evaluated with
 * Hairer's 7th-order continuous extension of the step that contains it and written to
 * out[c * stride + j].  The three extra stages (c14, c15, c16) do not feed back into y, so they
 * are only evaluated for steps that contain an output time. */
static int dop853_core(const cwo_potential *P, double x, double *y, double xend, double rtol, double atol,
                       double h, int64_t nmax, cwo_stats *st,
                       const double *tout, int64_t nout, double *out, int64_t stride)
{
    const int n = NEQ;
    int64_t jout = 0;
    double k1[NEQ], k2[NEQ], k3[NEQ], k4[NEQ], k5[NEQ], k6[NEQ], k7[NEQ], k8[NEQ], k9[NEQ], k10[NEQ];
    double yy1[NEQ];
    /* dop853(): defaults */
    if (nmax <= 0) nmax = 100000;
    const int64_t nstiff = 1;
    const double uround = 2.3E-16, safe = 0.9, fac1 = 0.333, fac2 = 6.0, beta = 0.0;
    double hmax = xend - x;
    /* dopcor() */
    double facold = 1.0E-4, expo1 = 1.0 / 8.0 - beta * 0.2, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    double posneg = sign_d(1.0, xend - x);
    double atoli = atol, rtoli = rtol;
    int last = 0, reject = 0, nonsti = 0, iasti = 0;
    int64_t nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0;
    double hlamb = 0.0, hnew, xph, err, err2, sk, erri, sqr, deno, fac, fac11;
    int i;

    fcn(P, x, y, k1);
    hmax = fabs(hmax);
    if (h == 0.0) return CWO_E_INCONSISTENT;   /* gala always passes dt0 != 0 (hinit not restated) */
    nfcn += 2;

    while (1) {
        if (nstep > nmax) { if (st) { st->n_accepted += naccpt; st->n_rejected += nrejct; st->n_fcn += nfcn; } return CWO_E_NMAX; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { if (st) { st->n_accepted += naccpt; st->n_rejected += nrejct; st->n_fcn += nfcn; } return CWO_E_STEP_TOO_SMALL; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;

        /* the twelve stages */
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * CW_A0201 * k1[i];
        fcn(P, x + CW_C02 * h, yy1, k2);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0301 * k1[i] + CW_A0302 * k2[i]);
        fcn(P, x + CW_C03 * h, yy1, k3);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0401 * k1[i] + CW_A0403 * k3[i]);
        fcn(P, x + CW_C04 * h, yy1, k4);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0501 * k1[i] + CW_A0503 * k3[i] + CW_A0504 * k4[i]);
        fcn(P, x + CW_C05 * h, yy1, k5);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0601 * k1[i] + CW_A0604 * k4[i] + CW_A0605 * k5[i]);
        fcn(P, x + CW_C06 * h, yy1, k6);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0701 * k1[i] + CW_A0704 * k4[i] + CW_A0705 * k5[i] + CW_A0706 * k6[i]);
        fcn(P, x + CW_C07 * h, yy1, k7);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0801 * k1[i] + CW_A0804 * k4[i] + CW_A0805 * k5[i] + CW_A0806 * k6[i] + CW_A0807 * k7[i]);
        fcn(P, x + CW_C08 * h, yy1, k8);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A0901 * k1[i] + CW_A0904 * k4[i] + CW_A0905 * k5[i] + CW_A0906 * k6[i] + CW_A0907 * k7[i] + CW_A0908 * k8[i]);
        fcn(P, x + CW_C09 * h, yy1, k9);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1001 * k1[i] + CW_A1004 * k4[i] + CW_A1005 * k5[i] + CW_A1006 * k6[i] + CW_A1007 * k7[i] + CW_A1008 * k8[i] + CW_A1009 * k9[i]);
        fcn(P, x + CW_C10 * h, yy1, k10);
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1101 * k1[i] + CW_A1104 * k4[i] + CW_A1105 * k5[i] + CW_A1106 * k6[i] + CW_A1107 * k7[i] + CW_A1108 * k8[i] + CW_A1109 * k9[i] + CW_A1110 * k10[i]);
        fcn(P, x + CW_C11 * h, yy1, k2);
        xph = x + h;
        for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1201 * k1[i] + CW_A1204 * k4[i] + CW_A1205 * k5[i] + CW_A1206 * k6[i] + CW_A1207 * k7[i] + CW_A1208 * k8[i] + CW_A1209 * k9[i] + CW_A1210 * k10[i] + CW_A1211 * k2[i]);
        fcn(P, xph, yy1, k3);
        nfcn += 11;
        for (i = 0; i < n; i++) {
            k4[i] = CW_B01 * k1[i] + CW_B06 * k6[i] + CW_B07 * k7[i] + CW_B08 * k8[i] + CW_B09 * k9[i] + CW_B10 * k10[i] + CW_B11 * k2[i] + CW_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }

        /* error estimation */
        err = 0.0; err2 = 0.0;
        for (i = 0; i < n; i++) {
            sk = atoli + rtoli * max_d(fabs(y[i]), fabs(k5[i]));
            erri = k4[i] - CW_BHH1 * k1[i] - CW_BHH2 * k9[i] - CW_BHH3 * k3[i];
            sqr = erri / sk;
            err2 += sqr * sqr;
            erri = CW_ER01 * k1[i] + CW_ER06 * k6[i] + CW_ER07 * k7[i] + CW_ER08 * k8[i] + CW_ER09 * k9[i] + CW_ER10 * k10[i] + CW_ER11 * k2[i] + CW_ER12 * k3[i];
            sqr = erri / sk;
            err += sqr * sqr;
        }
        deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (deno * (double)n));
        if (!(err == err) || isinf(err)) {  /* not in Hairer: NaN would spin to nmax; both end as a failed orbit */
            if (st) { st->n_accepted += naccpt; st->n_rejected += nrejct; st->n_fcn += nfcn; }
            return CWO_E_NONFINITE;
        }

        /* computation of hnew */
        fac11 = pow(err, expo1);
        fac = fac11 / pow(facold, beta);            /* Lund stabilisation (beta = 0 -> 1) */
        fac = max_d(facc2, min_d(facc1, fac / safe));
        hnew = h / fac;

        if (err <= 1.0) {
            /* step accepted */
            facold = max_d(err, 1.0E-4);
            naccpt++;
            fcn(P, xph, k5, k4);
            nfcn++;
            /* stiffness detection */
            if (!(naccpt % nstiff) || (iasti > 0)) {
                double stnum = 0.0, stden = 0.0;
                for (i = 0; i < n; i++) {
                    sqr = k4[i] - k3[i]; stnum += sqr * sqr;
                    sqr = k5[i] - yy1[i]; stden += sqr * sqr;
                }
                if (stden > 0.0) hlamb = h * sqrt(stnum / stden);
                if (hlamb > 6.1) {
                    nonsti = 0; iasti++;
                    if (iasti == 15) { if (st) { st->n_accepted += naccpt; st->n_rejected += nrejct; st->n_fcn += nfcn; } return CWO_E_STIFF; }
                } else {
                    nonsti++;
                    if (nonsti == 6) iasti = 0;
                }
            }
            /* dense output: dop853.c "final preparation for dense output" + contd8 */
            if (jout < nout && (tout[jout] - xph) * posneg <= 0.0) {
                double rc[8][NEQ], ka[NEQ], kb[NEQ], kc[NEQ];
                for (i = 0; i < n; i++) {
                    double ydiff = k5[i] - y[i];
                    double bspl = h * k1[i] - ydiff;
                    rc[0][i] = y[i];
                    rc[1][i] = ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = ydiff - h * k4[i] - bspl;
                    rc[4][i] = CW_D401 * k1[i] + CW_D406 * k6[i] + CW_D407 * k7[i] + CW_D408 * k8[i] + CW_D409 * k9[i] + CW_D410 * k10[i] + CW_D411 * k2[i] + CW_D412 * k3[i];
                    rc[5][i] = CW_D501 * k1[i] + CW_D506 * k6[i] + CW_D507 * k7[i] + CW_D508 * k8[i] + CW_D509 * k9[i] + CW_D510 * k10[i] + CW_D511 * k2[i] + CW_D512 * k3[i];
                    rc[6][i] = CW_D601 * k1[i] + CW_D606 * k6[i] + CW_D607 * k7[i] + CW_D608 * k8[i] + CW_D609 * k9[i] + CW_D610 * k10[i] + CW_D611 * k2[i] + CW_D612 * k3[i];
                    rc[7][i] = CW_D701 * k1[i] + CW_D706 * k6[i] + CW_D707 * k7[i] + CW_D708 * k8[i] + CW_D709 * k9[i] + CW_D710 * k10[i] + CW_D711 * k2[i] + CW_D712 * k3[i];
                }
                /* the next three function evaluations (k4 = f(x+h, y_new) is stage 13) */
                for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1401 * k1[i] + CW_A1407 * k7[i] + CW_A1408 * k8[i] + CW_A1409 * k9[i] + CW_A1410 * k10[i] + CW_A1411 * k2[i] + CW_A1412 * k3[i] + CW_A1413 * k4[i]);
                fcn(P, x + 0.1 * h, yy1, ka);
                for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1501 * k1[i] + CW_A1506 * k6[i] + CW_A1507 * k7[i] + CW_A1508 * k8[i] + CW_A1511 * k2[i] + CW_A1512 * k3[i] + CW_A1513 * k4[i] + CW_A1514 * ka[i]);
                fcn(P, x + 0.2 * h, yy1, kb);
                for (i = 0; i < n; i++) yy1[i] = y[i] + h * (CW_A1601 * k1[i] + CW_A1606 * k6[i] + CW_A1607 * k7[i] + CW_A1608 * k8[i] + CW_A1609 * k9[i] + CW_A1613 * k4[i] + CW_A1614 * ka[i] + CW_A1615 * kb[i]);
                fcn(P, x + 7. / 9. * h, yy1, kc);
                nfcn += 3;
                for (i = 0; i < n; i++) {
                    rc[4][i] = h * (rc[4][i] + CW_D413 * k4[i] + CW_D414 * ka[i] + CW_D415 * kb[i] + CW_D416 * kc[i]);
                    rc[5][i] = h * (rc[5][i] + CW_D513 * k4[i] + CW_D514 * ka[i] + CW_D515 * kb[i] + CW_D516 * kc[i]);
                    rc[6][i] = h * (rc[6][i] + CW_D613 * k4[i] + CW_D614 * ka[i] + CW_D615 * kb[i] + CW_D616 * kc[i]);
                    rc[7][i] = h * (rc[7][i] + CW_D713 * k4[i] + CW_D714 * ka[i] + CW_D715 * kb[i] + CW_D716 * kc[i]);
                }
                /* solout: every requested time inside (xold, x], contd8 */
                while (jout < nout && (tout[jout] - xph) * posneg <= 0.0) {
                    double s = (tout[jout] - x) / h, s1 = 1.0 - s;
                    for (i = 0; i < n; i++)
                        out[i * stride + jout] = rc[0][i] + s * (rc[1][i] + s1 * (rc[2][i] + s * (rc[3][i] + s1 * (rc[4][i] + s * (rc[5][i] + s1 * (rc[6][i] + s * rc[7][i]))))));
                    jout++;
                }
            }
            memcpy(k1, k4, n * sizeof(double));
            memcpy(y, k5, n * sizeof(double));
            x = xph;
            if (last) {
                if (st) { st->n_accepted += naccpt; st->n_rejected += nrejct; st->n_fcn += nfcn; }
                return 1;
            }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * min_d(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            /* step rejected */
            hnew = h / min_d(facc1, fac11 / safe);
            reject = 1;
            if (naccpt >= 1) nrejct = nrejct + 1;
            last = 0;
        }
        h = hnew;
    }
}

int cwo_dop853(const cwo_potential *P, double x, double *y, double xend, double rtol, double atol,
               double h, int64_t nmax, cwo_stats *st)
{
    return dop853_core(P, x, y, xend, rtol, atol, h, nmax, st, NULL, 0, NULL, 0);
}

/* one dop853() call with dense output at tout[0..nout-1] (exposed for the interpolant tests) */
int cwo_dop853_dense(const cwo_potential *P, double x, double *y, double xend, double rtol, double atol,
                     double h, int64_t nmax, cwo_stats *st, const double *tout, int64_t nout, double *out,
                     int64_t stride)
{
    return dop853_core(P, x, y, xend, rtol, atol, h, nmax, st, tout, nout, out, stride);
}

/* DENSE-OUTPUT variant of dop853_integrate_hamiltonian (newer gala releases
 * evaluate the requested times with Hairer's continuous extension instead of stopping the
 * integrator on every one): ONE dop853() call from t[0] to t[nt-1], h = t[1]-t[0],
 * hmax = t[nt-1]-t[0], nmax steps for the whole segment; interior samples by contd8, the last
 * sample is the integrator's own end state. */
int cwo_dop853_dense_segment(const cwo_potential *P, double *w, const double *t, int64_t nt,
                             double rtol, double atol, int64_t nmax, double *out, int64_t stride,
                             cwo_stats *st)
{
    if (out) for (int c = 0; c < 6; c++) out[c * stride] = w[c];
    if (nt < 2) return CWO_OK;
    int res = dop853_core(P, t[0], w, t[nt - 1], rtol, atol, t[1] - t[0], nmax, st,
                          out ? t + 1 : NULL, out ? nt - 2 : 0, out ? out + 1 : NULL, stride);
    if (res < 0) return res;
    if (out) for (int c = 0; c < 6; c++) out[c * stride + nt - 1] = w[c];
    return CWO_OK;
}

/* dop853.pyx dop853_integrate_hamiltonian for one orbit over t[0..nt-1] (Myr):
 * dt0 = t[1]-t[0]; a fresh dop853(x=t[j-1], xend=t[j], h=dt0) per output interval. */
int cwo_dop853_segment(const cwo_potential *P, double *w, const double *t, int64_t nt,
                       double rtol, double atol, int64_t nmax, double *out, int64_t stride,
                       cwo_stats *st)
{
    if (out) for (int c = 0; c < 6; c++) out[c * stride] = w[c];
    if (nt < 2) return CWO_OK;
    double dt0 = t[1] - t[0];
    for (int64_t j = 1; j < nt; j++) {
        int res = cwo_dop853(P, t[j - 1], w, t[j], rtol, atol, dt0, nmax, st);
        if (res < 0) return res;
        if (out) for (int c = 0; c < 6; c++) out[c * stride + j] = w[c];
    }
    return CWO_OK;
}

/* ------------------------------------------------------------------ the kicks.py driver */

/* potential.integrate_orbit(w, t=T[first..last]) in galactic units */
static int integrate_segment(const cwo_potential *P, int integrator, double rtol, double atol,
                             int64_t nmax, double *w, const double *tmyr, int64_t nt,
                             double *out, int64_t stride, cwo_stats *st)
{
    if (integrator == CWO_LEAPFROG) { cwo_leapfrog(P, w, tmyr, nt, out, stride); return CWO_OK; }
    if (integrator == CWO_DOP853_DENSE) return cwo_dop853_dense_segment(P, w, tmyr, nt, rtol, atol, nmax, out, stride, st);
    return cwo_dop853_segment(P, w, tmyr, nt, rtol, atol, nmax, out, stride, st);
}

/*
 * One orbit of integrate_orbit_with_events (kicks.py:97-209), one attempt (the retry loop
 * kicks.py:113,193-199 is host logic in the caller).
 *
 *   t1,t2,dt,ev_time : in the orbit's GRID unit  (seconds on the events branch kicks.py:118,
 *                      Myr on the no-events branch kicks.py:100-104, where gala builds the grid)
 *   to_myr           : grid unit -> Myr factor (astropy multiplies by 1/3.15576e13; 1.0 for Myr)
 *   flags & 1        : append t2 as the last node (kicks.py:121-122)
 *   ev_dv            : n_ev x 3, already rotated (get_kick_differential) and in kpc/Myr
 *   n_nodes_expected : trajectory capacity when store_all (samples written to
 *                      traj[c*stride + j], times to traj_t[j]); ignored otherwise
 * Outputs: w_final[6]; ev_step[e] = grid node at which kick e is applied; *n_nodes.
 */
int cwo_integrate_orbit(const cwo_potential *P, int integrator, double rtol, double atol, int64_t nmax,
                        const double *w0, double t1, double t2, double dt, double to_myr, int flags,
                        int64_t n_ev, const double *ev_time, const double *ev_dv,
                        int store_all, double *traj, int64_t stride, double *traj_t, int64_t cap,
                        double *w_final, int32_t *ev_step, int64_t *n_nodes, cwo_stats *st)
{
    int64_t n = cwo_time_grid(t1, t2, dt, flags & CWO_FLAG_APPEND_T2, NULL, 0);
    if (n_nodes) *n_nodes = n;
    if (n <= 0) return CWO_E_INCONSISTENT;
    if (dt < 0.0 && n_ev > 0) return CWO_E_INCONSISTENT;   /* kicks.py only ever integrates forward */
    if (store_all && cap < n) return CWO_E_CAPACITY;
    double *T = (double *)malloc(sizeof(double) * (size_t)n * 2);
    if (!T) return CWO_E_CAPACITY;
    double *Tm = T + n;
    cwo_time_grid(t1, t2, dt, flags & CWO_FLAG_APPEND_T2, T, n);
    for (int64_t j = 0; j < n; j++) Tm[j] = T[j] * to_myr;     /* Quantity.decompose(galactic) */

    double w[6];
    memcpy(w, w0, sizeof w);
    int status = CWO_OK;
    int64_t n_written = 0;             /* samples appended to orbit_data so far */
    double cursor = T[0];              /* kicks.py:125 */

    for (int64_t e = 0; e < n_ev && status == CWO_OK; e++) {
        /* timestep_mask = (timesteps >= time_cursor) & (timesteps < t1 + tphys)   kicks.py:134 */
        int64_t first = -1, last = -1;
        for (int64_t j = 0; j < n; j++) {
            if (T[j] >= cursor && T[j] < ev_time[e]) { if (first < 0) first = j; last = j; }
        }
        if (first >= 0) {
            int64_t nt = last - first + 1;
            double *out = NULL;
            if (store_all) out = traj + n_written;   /* writes nt samples; only nt-1 are kept */
            status = integrate_segment(P, integrator, rtol, atol, nmax, w, Tm + first, nt,
                                       out, stride, st);
            n_written += nt - 1;                 /* orbit.data[:-1]          kicks.py:148 */
            cursor = T[last];                    /* kicks.py:154 */
            if (ev_step) ev_step[e] = (int32_t)last;
        } else {
            cursor = ev_time[e];                 /* kicks.py:157 */
            /* the kick lands on the node the state currently sits at */
            if (ev_step) ev_step[e] = (int32_t)n_written;
        }
        /* current_w0.vel + kick_differential     kicks.py:160-172 */
        w[3] += ev_dv[3 * e + 0];
        w[4] += ev_dv[3 * e + 1];
        w[5] += ev_dv[3 * e + 2];
    }
    if (status == CWO_OK && (n_ev == 0 || cursor < T[n - 1])) {       /* kicks.py:175 / :102 */
        int64_t first = -1;
        for (int64_t j = 0; j < n; j++) if (T[j] >= cursor) { first = j; break; }
        if (first >= 0) {
            int64_t nt = n - first;
            double *out = NULL;
            if (store_all) out = traj + n_written;
            status = integrate_segment(P, integrator, rtol, atol, nmax, w, Tm + first, nt,
                                       out, stride, st);
            n_written += nt;
        }
    } else if (status == CWO_OK && store_all) {
        /* no tail: the state after the last kick is never stored by the reference */
    }
    /* Orbit(pos, vel, t=timesteps) requires len(data) == len(timesteps)   kicks.py:187-189 */
    if (status == CWO_OK && n_written != n) status = CWO_E_INCONSISTENT;
    if (status == CWO_OK && store_all && traj_t) for (int64_t j = 0; j < n; j++) traj_t[j] = Tm[j];
    memcpy(w_final, w, sizeof w);
    free(T);
    return status;
}

/*
 * Batch driver = the work Population.perform_galactic_evolution fans out over its pool
 * (pop.py:1240-1266), here a pthread pool pulling chunks of orbits from an atomic counter.
 * SoA layouts identical to the C-ABI of the product (include/cogsworth_b200.h) so the parity
 * tests feed both the same buffers.
 *   w0 / w_final : 6 x n (component-major);  ev_off : n+1 CSR offsets into ev_time / ev_dv(K x 3)
 *   traj_off     : n+1 offsets, traj_pos/vel : 3 x total, traj_t : total   (store_all only)
 */
typedef struct {
    const cwo_potential *P;
    int integrator; double rtol, atol; int64_t nmax;
    int64_t n_orbits;
    const double *w0, *t1, *t2, *dt, *to_myr; const int32_t *flags;
    const int64_t *ev_off; const double *ev_time, *ev_dv;
    double *w_final; int32_t *status, *n_nodes_out, *ev_step;
    const int64_t *traj_off; double *traj_pos, *traj_vel, *traj_t;
    int64_t next;                  /* work counter (atomic) */
    int64_t acc, rej, nf;          /* stats (atomic adds) */
} batch_job;

static void batch_one(batch_job *J, int64_t i, cwo_stats *tot)
{
    const int64_t n_orbits = J->n_orbits;
    int store_all = J->traj_off != NULL;
    int64_t total = store_all ? J->traj_off[n_orbits] : 0;
    double w[6], wf[6];
    for (int c = 0; c < 6; c++) w[c] = J->w0[c * n_orbits + i];
    int64_t e0 = J->ev_off ? J->ev_off[i] : 0, e1 = J->ev_off ? J->ev_off[i + 1] : 0;
    int64_t nn = 0;
    cwo_stats st = {0, 0, 0};
    double *tmp = NULL;
    int64_t cap = 0;
    if (store_all) {
        cap = J->traj_off[i + 1] - J->traj_off[i];
        tmp = (double *)malloc(sizeof(double) * 6 * (size_t)(cap + 1));
    }
    int s = cwo_integrate_orbit(J->P, J->integrator, J->rtol, J->atol, J->nmax, w, J->t1[i], J->t2[i],
                                J->dt[i], J->to_myr[i], J->flags[i], e1 - e0,
                                J->ev_time ? J->ev_time + e0 : NULL, J->ev_dv ? J->ev_dv + 3 * e0 : NULL,
                                store_all, tmp, cap + 1,
                                (store_all && J->traj_t) ? J->traj_t + J->traj_off[i] : NULL, cap,
                                wf, J->ev_step ? J->ev_step + e0 : NULL, &nn, &st);
    if (store_all && tmp) {
        if (s == CWO_OK) {
            for (int c = 0; c < 3; c++)
                for (int64_t j = 0; j < nn; j++) {
                    J->traj_pos[c * total + J->traj_off[i] + j] = tmp[c * (cap + 1) + j];
                    J->traj_vel[c * total + J->traj_off[i] + j] = tmp[(3 + c) * (cap + 1) + j];
                }
        }
        free(tmp);
    }
    for (int c = 0; c < 6; c++) J->w_final[c * n_orbits + i] = wf[c];
    if (J->status) J->status[i] = s;
    if (J->n_nodes_out) J->n_nodes_out[i] = (int32_t)nn;
    tot->n_accepted += st.n_accepted; tot->n_rejected += st.n_rejected; tot->n_fcn += st.n_fcn;
}

static void *batch_worker(void *arg)
{
    batch_job *J = (batch_job *)arg;
    cwo_stats tot = {0, 0, 0};
    const int64_t chunk = 16;
    for (;;) {
        int64_t b = __atomic_fetch_add(&J->next, chunk, __ATOMIC_RELAXED);
        if (b >= J->n_orbits) break;
        int64_t e = b + chunk < J->n_orbits ? b + chunk : J->n_orbits;
        for (int64_t i = b; i < e; i++) batch_one(J, i, &tot);
    }
    __atomic_fetch_add(&J->acc, tot.n_accepted, __ATOMIC_RELAXED);
    __atomic_fetch_add(&J->rej, tot.n_rejected, __ATOMIC_RELAXED);
    __atomic_fetch_add(&J->nf, tot.n_fcn, __ATOMIC_RELAXED);
    return NULL;
}

int cwo_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int cwo_integrate_batch(const cwo_potential *P, int integrator, double rtol, double atol, int64_t nmax,
                        int64_t n_orbits, const double *w0, const double *t1, const double *t2,
                        const double *dt, const double *to_myr, const int32_t *flags,
                        const int64_t *ev_off, const double *ev_time, const double *ev_dv,
                        double *w_final, int32_t *status, int32_t *n_nodes_out, int32_t *ev_step,
                        const int64_t *traj_off, double *traj_pos, double *traj_vel, double *traj_t,
                        int64_t *stats3, int n_threads)
{
    batch_job J = {P, integrator, rtol, atol, nmax, n_orbits, w0, t1, t2, dt, to_myr, flags,
                   ev_off, ev_time, ev_dv, w_final, status, n_nodes_out, ev_step,
                   traj_off, traj_pos, traj_vel, traj_t, 0, 0, 0, 0};
    if (n_threads <= 0) n_threads = cwo_max_threads();
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) {
        batch_worker(&J);
    } else {
        pthread_t th[256];
        int started = 0;
        for (int k = 0; k < n_threads; k++)
            if (pthread_create(&th[started], NULL, batch_worker, &J) == 0) started++;
        if (started == 0) batch_worker(&J);
        for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
    }
    if (stats3) { stats3[0] = J.acc; stats3[1] = J.rej; stats3[2] = J.nf; }
    return 0;
}
