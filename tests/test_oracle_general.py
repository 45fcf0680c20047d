"""CPU checks of the oracle's general potentials (rotated / non-diagonal components of the web UI's
menu -- reference online-interface/potentials.py:93-131 -- and gp.TimeInterpolatedPotential,
docs/tutorials/potentials/evolving_potentials.ipynb cells 5 and 13) against independent restatements:
finite differences of the potential values, scipy's natural cubic spline, scipy's incomplete gamma function,
scipy's DOP853 with a time-dependent right-hand side, and the closed form of Law & Majewski (2010) eq. 3.
"""
import numpy as np
import pytest

from cogsworth_b200 import potential as P
from cogsworth_b200.constants import G_GALACTIC


def _fd_gradient(O, pot, q, h=1e-5):
    g = np.zeros(3)
    for k in range(3):
        e = np.zeros(3); e[k] = h
        g[k] = (O.value(pot, q + e) - O.value(pot, q - e)) / (2 * h)
    return g


@pytest.mark.parametrize("pot", [
    P.LongMuraliBarPotential(1e10, 10.0, 3.0, 1.0),
    P.LongMuraliBarPotential(2e10, 4.0, 0.5, 0.3, alpha=0.4363),
    P.BovyMWPotential2014(),
    P.LM10Potential(),
], ids=["bar", "bar-rotated", "bovy2014", "lm10"])
def test_gradient_is_the_derivative_of_the_value(oracle, pot):
    rng = np.random.default_rng(5)
    for _ in range(40):
        q = rng.normal(size=3) * np.array([8.0, 8.0, 3.0])
        g = oracle.gradient(pot, q)
        fd = _fd_gradient(oracle, pot, q)
        assert np.allclose(g, fd, rtol=2e-7, atol=1e-9 * np.abs(g).max()), (q, g, fd)


@pytest.mark.parametrize("par", [(0.2, 20.0, 1.0, 0.9, 0.8), (0.15, 8.0, 1.0, 0.6, 0.4)])
def test_leesuto_gradient_is_the_derivative_of_lee_suto_2003(oracle, par):
    # F2 and F3 cancel badly at small u in double precision, so the derivative is taken with 40 digits
    import mpmath as mp
    mp.mp.dps = 40
    v_c, r_s, a, b, c = par
    pot = P.LeeSutoTriaxialNFWPotential(*par)

    def phi(x, y, z):
        eb2, ec2 = 1 - mp.mpf(b / a) ** 2, 1 - mp.mpf(c / a) ** 2
        phi0 = mp.mpf(v_c) ** 2 / (mp.log(2) - mp.mpf(1) / 2 + (mp.log(2) - mp.mpf(3) / 4) * (eb2 + ec2))
        r = mp.sqrt(x * x + y * y + z * z); u = r / r_s; L = mp.log(1 + u)
        F1 = -L / u
        F2 = -mp.mpf(1) / 3 + (2 * u * u - 3 * u + 6) / (6 * u * u) + (1 / u - 1 / u ** 3) * L
        F3 = (u * u - 3 * u - 6) / (2 * u * u * (1 + u)) + 3 * L / u ** 3
        return phi0 * (F1 + (eb2 + ec2) / 2 * F2 + (eb2 * y * y + ec2 * z * z) / (2 * r * r) * F3)

    rng = np.random.default_rng(7)
    for _ in range(25):
        q = rng.normal(size=3) * np.array([15.0, 15.0, 8.0])
        want = [float(mp.diff(phi, tuple(mp.mpf(v) for v in q), n=tuple(int(k == i) for k in range(3)))) for i in range(3)]
        g = oracle.gradient(pot, q)
        assert np.allclose(g, want, rtol=1e-9, atol=1e-11 * np.abs(g).max()), (q, g, want)
        assert np.isclose(oracle.value(pot, q), float(phi(*[mp.mpf(v) for v in q])), rtol=1e-10)


def test_lm10_halo_is_law_majewski_eq3(oracle):
    # Phi_halo = v_halo^2 ln(C1 x^2 + C2 y^2 + C3 x y + (z / q_z)^2 + r_halo^2), v_c^2 / 2 = v_halo^2
    pot = P.LM10Potential()
    v_c, r_h, q1, q2, qz = pot.params[2]
    phi = np.deg2rad(97.0)
    C1 = np.cos(phi) ** 2 / q1 ** 2 + np.sin(phi) ** 2 / q2 ** 2
    C2 = np.cos(phi) ** 2 / q2 ** 2 + np.sin(phi) ** 2 / q1 ** 2
    C3 = 2 * np.sin(phi) * np.cos(phi) * (1 / q1 ** 2 - 1 / q2 ** 2)
    halo = P.CompositePotential(names=["halo"], types=[pot.types[2]], params=[pot.params[2]],
                                rotations=[pot.rotations[2]])
    rng = np.random.default_rng(6)
    for _ in range(20):
        x, y, z = rng.normal(size=3) * 20
        S = C1 * x * x + C2 * y * y + C3 * x * y + (z / qz) ** 2 + r_h ** 2
        assert np.isclose(oracle.value(halo, np.array([x, y, z])), 0.5 * v_c ** 2 * np.log(S), rtol=1e-13)
        g = oracle.gradient(halo, np.array([x, y, z]))
        want = v_c ** 2 / S * np.array([C1 * x + 0.5 * C3 * y, C2 * y + 0.5 * C3 * x, z / qz ** 2])
        assert np.allclose(g, want, rtol=1e-12)
    assert np.isclose(v_c * 977.7922216807891, np.sqrt(2) * 121.858)


def test_powerlaw_cutoff_enclosed_mass_matches_scipy(oracle):
    from scipy.special import gammainc
    m, alpha, r_c = 4501365375.06545, 1.8, 1.9
    pot = P.CompositePotential(names=["b"], types=[P.COMP_POWERLAW_CUTOFF], params=[(m, alpha, r_c)])
    for r in [1e-3, 0.05, 0.5, 1.0, 1.9, 2.7, 5.0, 12.0, 40.0]:
        g = oracle.gradient(pot, np.array([r, 0.0, 0.0]))
        want = G_GALACTIC * m * gammainc(1.5 - alpha / 2, (r / r_c) ** 2) / r ** 2
        assert np.isclose(g[0], want, rtol=5e-14), (r, g[0], want)
    # total mass m: Kepler far outside the cutoff, and the value vanishes at infinity
    assert np.isclose(oracle.value(pot, np.array([0.0, 300.0, 0.0])), -G_GALACTIC * m / 300.0, rtol=1e-12)


def test_bovy2014_is_normalised_like_galpy(oracle):
    # MWPotential2014 is built so that v_circ(8 kpc) = 220 km/s with the disk giving 60 %, the halo 35 %, the bulge 5 %
    pot = P.BovyMWPotential2014()
    q = np.array([8.0, 0.0, 0.0])
    vc = oracle.circular_velocity(pot, q) * 977.7922216807891
    assert abs(vc - 220.0) < 0.05
    fr = []
    for i in range(3):
        one = P.CompositePotential(names=[pot.names[i]], types=[pot.types[i]], params=[pot.params[i]])
        fr.append(oracle.gradient(one, q)[0])
    fr = np.array(fr) / sum(fr)
    assert np.allclose(fr, [0.60, 0.05, 0.35], atol=2e-3)


@pytest.mark.parametrize("interp", ["cubic", "linear"])
def test_time_interpolation_matches_scipy(oracle, interp):
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(8)
    t = np.sort(rng.uniform(0, 12000, 9)); t[0], t[-1] = 0.0, 12000.0
    m = 1e11 * (1 + rng.uniform(0, 9, 9))
    base = P.CompositePotential(names=["h"], types=[P.COMP_NFW_SPHERICAL], params=[(1e12, 10.0, 0.0)])
    pot = P.time_interpolated(base, t, [m], interp=interp)
    q = np.array([8.0, 1.0, -2.0])
    unit = oracle.gradient(P.CompositePotential(names=["h"], types=[P.COMP_NFW_SPHERICAL], params=[(1.0, 10.0, 0.0)]), q)
    ref = CubicSpline(t, m, bc_type="natural") if interp == "cubic" else (lambda x: np.interp(x, t, m))
    for tt in np.concatenate([rng.uniform(0, 12000, 50), t, [-5.0, 12500.0]]):
        g = oracle.gradient_t(pot, tt, q)
        want = float(ref(np.clip(tt, 0.0, 12000.0)))          # the end values are held outside the knots
        assert np.allclose(g, unit * want, rtol=1e-12), (tt, g / unit, want)


def test_linear_masses_interpolate_linearly_whatever_the_method(oracle):
    # the tutorial's case: masses on a linspace -- the natural spline through collinear points is the line
    t = np.linspace(0, 12000, 20)
    m = np.linspace(1e11, 1e12, 20)
    base = P.CompositePotential(names=["h"], types=[P.COMP_NFW_SPHERICAL], params=[(1e12, 10.0, 0.0)])
    q = np.array([3.0, -4.0, 1.0])
    gc = oracle.gradient_t(P.time_interpolated(base, t, [m], "cubic"), 4321.0, q)
    gl = oracle.gradient_t(P.time_interpolated(base, t, [m], "linear"), 4321.0, q)
    assert np.allclose(gc, gl, rtol=1e-13)
    static = oracle.gradient(base, q)
    assert np.allclose(gc, static * (1e11 + 9e11 * 4321.0 / 12000.0) / 1e12, rtol=1e-13)


def _rhs(O, pot):
    def f(t, y):
        g = O.gradient_t(pot, t, y[:3])
        return np.concatenate([y[3:], -g])
    return f


def test_dop853_in_an_evolving_potential_matches_scipy(oracle):
    """The tutorial's test particle (cell 7) in the growing NFW halo (cell 5): the oracle's DOP853 with the clock threaded
    through every stage against scipy's DOP853 of the same time-dependent right-hand side."""
    from scipy.integrate import solve_ivp
    t = np.linspace(0, 12000, 20)
    base = P.CompositePotential(names=["h"], types=[P.COMP_NFW_SPHERICAL], params=[(1e12, 10.0, 0.0)])
    pot = P.time_interpolated(base, t, [np.linspace(1e11, 1e12, 20)])
    w0 = np.array([8.0, 0.0, 0.0, 50.0 / 977.7922216807891, 50.0 / 977.7922216807891, -5.0 / 977.7922216807891])
    T = 3000.0
    r = oracle.integrate_orbit(pot, w0, 0.0, T, 1.0, integrator=oracle.DOP853, rtol=1e-11, atol=1e-11)
    assert r["status"] == 0
    w = r["w_final"]
    sol = solve_ivp(_rhs(oracle, pot), (0.0, w_end_time(oracle, T)), w0, method="DOP853", rtol=1e-12, atol=1e-14)
    assert np.allclose(w, sol.y[:, -1], rtol=2e-8, atol=1e-9)
    # ... and the evolution matters: the static halo gives a different orbit
    ws = oracle.integrate_orbit(base, w0, 0.0, T, 1.0, integrator=oracle.DOP853, rtol=1e-11, atol=1e-11)["w_final"]
    assert np.abs(ws[:3] - w[:3]).max() > 0.5
    # dense-output form: same orbit within the DOPRI853 tolerance
    rd = oracle.integrate_orbit(pot, w0, 0.0, T, 1.0, integrator=oracle.DOP853_DENSE, rtol=1e-11, atol=1e-11)
    assert rd["status"] == 0 and np.allclose(rd["w_final"], w, rtol=5e-7, atol=1e-8)   # a near-radial orbit


def w_end_time(oracle, T):
    """gala's own grid (no appended t2) ends on the last node < T."""
    n = oracle.integrate_orbit(P.MilkyWayPotential("v1"), np.array([8.0, 0, 0, 0, 0.2, 0]), 0.0, T, 1.0)["n_nodes"]
    return float(n - 1)


def test_leapfrog_in_an_evolving_potential_uses_the_new_nodes_time(oracle):
    pot = P.evolving_milky_way(np.linspace(0, 12000, 5), np.linspace(0.1, 1.0, 5))
    w0 = np.array([8.0, 0.3, 0.1, 0.01, 0.21, 0.005])
    tg = 100.0 + np.arange(301.0)
    traj = oracle.leapfrog(pot, w0, tg, store_all=True)
    # literal KDK with the gradient of step j-1 -> j taken at t[j]
    x, v = w0[:3].copy(), w0[3:].copy()
    dt = tg[1] - tg[0]
    vh = v - oracle.gradient_t(pot, tg[0], x) * dt / 2
    for j in range(1, len(tg)):
        x = x + vh * dt
        g = oracle.gradient_t(pot, tg[j], x)
        v = vh - g * dt / 2
        vh = vh - g * dt
    got = np.asarray(traj).reshape(6, -1)[:, -1]
    assert np.allclose(got, np.concatenate([x, v]), rtol=1e-13, atol=1e-15)
