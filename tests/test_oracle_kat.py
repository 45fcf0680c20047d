"""CPU tests of the oracle (oracle/cw_oracle.c) against every known answer the reference holds for
this path: the executed-notebook values in tests/golden/gala_notebook_kats.json
(extracted by tools/extract_notebook_kats.py) and scipy's compiled Hairer DOP853."""
import json
import os

import numpy as np
import pytest

from cogsworth_b200 import constants as K
from cogsworth_b200.potential import MilkyWayPotential, mn3_exponential_disk

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "gala_notebook_kats.json")))


def test_kat1_circular_velocity(oracle):
    v2 = MilkyWayPotential("v2")
    vc = oracle.circular_velocity(v2, [8.5, 0.0, 0.0]) * K.KMS_PER_KPC_MYR
    # all 11 printed digits of the notebook value
    assert abs(vc - KATS["KAT1_vcirc_kms_at_8p5kpc_v2"]) < 5e-9
    # the KAT discriminates the parameter set: v1 is 230.75 km/s
    v1 = oracle.circular_velocity(MilkyWayPotential("v1"), [8.5, 0.0, 0.0]) * K.KMS_PER_KPC_MYR
    assert abs(v1 - 230.748) < 1e-2 and abs(v1 - vc) > 1.0


def test_kat2_structure():
    v2 = MilkyWayPotential("v2")
    assert KATS["KAT2_component_order"] == ["disk", "bulge", "nucleus", "halo"]
    assert [n.split(".")[0] for n in v2.names] == ["disk"] * 3 + ["bulge", "nucleus", "halo"]
    assert "m=4.77e+10" in KATS["KAT2_disk_repr"] and "h_R=2.60" in KATS["KAT2_disk_repr"]
    terms = mn3_exponential_disk(4.7717e10, 2.6, 0.3)
    assert np.allclose([t[0] for t in terms], [7.8723070e9, -2.7562522e11, 3.2061842e11], rtol=1e-7)
    assert np.allclose([t[1] for t in terms], [1.5259432, 6.7827644, 5.8947996], rtol=1e-7)
    assert all(abs(t[2] - 0.2066374) < 1e-7 for t in terms)
    # nominal component masses sum to the 6.09e11 printed at docs/tutorials/pop_settings/potentials.ipynb:132
    assert abs((4.7717e10 + sum(p[0] for p in v2.params[3:])) / 6.09e11 - 1) < 1e-3


def test_kat4_time_grid(oracle):
    """Kicked-orbit grid: accumulate dt in seconds, append t2 (kicks.py:118-122)."""
    t_first, t_last = KATS["KAT4_t_first3_myr"], KATS["KAT4_t_last3_myr"]
    T = oracle.time_grid(t_first[0] * 1e-3 * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR,
                         KATS["KAT4_timestep_myr"] * K.SEC_PER_MYR, append_t2=True)
    assert len(T) == KATS["KAT4_n_samples"] == 77554
    tm = T * K.MYR_PER_SEC
    assert np.allclose(tm[:3], t_first, atol=1e-8, rtol=0)
    assert np.allclose(tm[-3:], t_last, atol=2e-8, rtol=0)
    assert tm[-1] == 12000.0
    # without the append the grid stops at the last node < t2 (gala's own rule, reference quirk Q1)
    T2 = oracle.time_grid(t_first[0], 12000.0, 0.1, append_t2=False)
    assert len(T2) == 77553 and T2[-1] < 12000.0


def test_kat5_dop853_samples(oracle):
    """Six printed phase-space samples of a real gala DOPRI853 orbit in MilkyWayPotential2022."""
    v2 = MilkyWayPotential("v2")
    for pos, vel, t in ((KATS["KAT5_pos_first3_kpc"], KATS["KAT5_vel_first3_kpc_per_myr"], KATS["KAT4_t_first3_myr"]),
                        (KATS["KAT5_pos_last3_kpc"], KATS["KAT5_vel_last3_kpc_per_myr"], KATS["KAT4_t_last3_myr"])):
        w = np.array(pos[0] + vel[0])
        import ctypes as C
        P = oracle.make_potential(v2)
        tt = np.array(t)
        out = np.empty((6, 3))
        dp = C.POINTER(C.c_double)
        rc = oracle.lib().cwo_dop853_segment(C.byref(P), w.ctypes.data_as(dp), tt.ctypes.data_as(dp), 3, 1e-10,
                                             1e-10, 0, out.ctypes.data_as(dp), 3, None)
        assert rc == 0
        got = out.T[1:]
        exp = np.array([pos[1] + vel[1], pos[2] + vel[2]])
        # the start sample is itself rounded to 8 decimals; positions move by v*dt with dt <= 0.2 Myr
        assert np.abs(got - exp).max() < 2.5e-8, np.abs(got - exp).max()
    # the last interval is the short appended one (0.0588 Myr): the position increment shrinks with it
    d1 = np.array(KATS["KAT5_pos_last3_kpc"][1]) - np.array(KATS["KAT5_pos_last3_kpc"][0])
    d2 = np.array(KATS["KAT5_pos_last3_kpc"][2]) - np.array(KATS["KAT5_pos_last3_kpc"][1])
    frac = (KATS["KAT4_t_last3_myr"][2] - KATS["KAT4_t_last3_myr"][1]) / 0.1
    assert abs(d2[0] / d1[0] - frac) < 1e-3


def test_dop853_matches_scipy_hairer(oracle):
    """scipy.integrate.ode('dop853') is the compiled Hairer code with the same controller: identical
    step sequences, results equal to rounding."""
    from scipy.integrate import ode
    v2 = MilkyWayPotential("v2")

    def f(t, y):
        g = oracle.gradient(v2, y[:3])
        return np.concatenate([y[3:], -g])

    rng = np.random.default_rng(5)
    for trial in range(6):
        y0 = np.array([rng.uniform(0.3, 12), rng.uniform(-3, 3), rng.uniform(-1, 1),
                       rng.uniform(-0.1, 0.1), rng.uniform(0.05, 0.25), rng.uniform(-0.05, 0.05)])
        y_or = y0.copy()
        y_sp = y0.copy()
        n_steps_total = 0
        for j in range(40):
            t0, t1 = 100.0 + j * 1.0, 101.0 + j * 1.0
            res, y_or, st = oracle.dop853(v2, t0, y_or, t1, rtol=1e-10, atol=1e-10, h=1.0)
            assert res == 1
            n_steps_total += st[0]
            r = ode(f).set_integrator("dop853", rtol=1e-10, atol=1e-10, first_step=1.0, nsteps=100000,
                                      safety=0.9, ifactor=6.0, dfactor=0.333, beta=0.0)
            r.set_initial_value(y_sp, t0)
            y_sp = r.integrate(t1)
            assert r.successful()
        rel = np.abs(y_or - y_sp) / np.maximum(np.abs(y_sp), 1e-3)
        assert rel.max() < 1e-11, (trial, rel.max(), n_steps_total)


def test_dop853_tableau_consistency():
    """The generated header reproduces scipy's tableau: row sums = c_i, sum b = 1, error weights sum to 0."""
    import re
    txt = open(os.path.join(HERE, "..", "include", "cw_dop853_coeffs.h")).read()
    vals = {m.group(1): float(m.group(2)) for m in re.finditer(r"#define CW_(\w+) ([-\d.e+]+)", txt)}
    for i in range(2, 13):
        row = sum(v for k, v in vals.items() if re.fullmatch(rf"A{i:02d}\d\d", k))
        assert abs(row - vals[f"C{i:02d}"]) < 1e-14, i
    assert abs(sum(v for k, v in vals.items() if re.fullmatch(r"B\d\d", k)) - 1) < 1e-14
    assert abs(sum(v for k, v in vals.items() if re.fullmatch(r"ER\d\d", k))) < 1e-14
    assert abs(vals["BHH1"] + vals["BHH2"] + vals["BHH3"] - 1) < 1e-14


def test_gradient_against_numpy_and_finite_difference(oracle):
    rng = np.random.default_rng(1)
    for ver in ("v1", "v2"):
        pot = MilkyWayPotential(ver)
        for _ in range(50):
            q = rng.normal(0, 6, 3)
            g = oracle.gradient(pot, q)
            # central differences of the potential value
            h = 1e-5
            fd = np.array([(oracle.value(pot, q + h * e) - oracle.value(pot, q - h * e)) / (2 * h) for e in np.eye(3)])
            assert np.allclose(g, fd, rtol=2e-6, atol=1e-12)


def test_extra_components_gradient_is_derivative_of_value(oracle):
    """Kepler / Plummer / Isochrone / Jaffe / axisymmetric Logarithmic: the restated gradients
    are the derivatives of the restated potential values, alone and inside a composite."""
    from cogsworth_b200.potential import CompositePotential
    singles = [(4, (3e10, 0.0, 0.0)), (5, (2e10, 1.3, 0.0)), (6, (5e10, 2.1, 0.0)), (7, (4e10, 3.0, 0.0)),
               (8, (0.2, 1.5, 0.8)), (9, (3e10, 2.5, 0.0)), (10, (2e7, 9.0, 0.0))]
    pots = [CompositePotential(names=["c"], types=[t], params=[p]) for t, p in singles]
    pots.append(CompositePotential(names=["tnfw"], types=[11], params=[(6.09e11, 10.0, 0.25, 1.0, 1.0)]))   # the tutorial's NFW
    pots.append(CompositePotential(names=["tnfw2"], types=[11], params=[(5e11, 15.0, 1.0, 0.8, 0.6)]))
    pots.append(CompositePotential(names=["tlog"], types=[12], params=[(0.18, 12.0, 1.38, 1.0, 1.36)]))     # LM10-like axis ratios
    pots.append(CompositePotential(names=list("abcdefghi"), types=[1, 4, 5, 6, 7, 8, 9, 10, 3],
                                   params=[(5e10, 3.0, 0.3)] + [p for _, p in singles] + [(5e11, 16.0, 0.0)]))
    rng = np.random.default_rng(3)
    for pot in pots:
        for _ in range(30):
            q = rng.normal(0, 5, 3)
            g = oracle.gradient(pot, q)
            h = 1e-5
            fd = np.array([(oracle.value(pot, q + h * e) - oracle.value(pot, q - h * e)) / (2 * h) for e in np.eye(3)])
            assert np.allclose(g, fd, rtol=5e-6, atol=1e-12), (pot.types, q, g, fd)
    # closed forms: Kepler circular speed, Plummer at r = 0 scale, logarithmic flat rotation curve
    vk = oracle.circular_velocity(pots[0], [4.0, 0.0, 0.0])
    assert abs(vk - np.sqrt(pots[0].G * 3e10 / 4.0)) < 1e-15
    vl = oracle.circular_velocity(pots[4], [400.0, 0.0, 0.0])
    assert abs(vl - 0.2) < 1e-5
    # Burkert: enclosed mass pi rho r0^3 (ln(1+x^2) + 2 ln(1+x) - 2 atan x) at x = 1; Kuzmin is even in z
    rho, r0 = 2e7, 9.0
    m_enc = np.pi * rho * r0 ** 3 * (np.log(2.0) + 2 * np.log(2.0) - 2 * np.arctan(1.0))
    vb = oracle.circular_velocity(pots[6], [r0, 0.0, 0.0])
    assert abs(vb - np.sqrt(pots[6].G * m_enc / r0)) < 1e-15
    # a flattened NFW with a = b = c = 1 is the spherical one, bit for bit
    sph = CompositePotential(names=["h"], types=[3], params=[(5e11, 15.0, 0.0)])
    tri = CompositePotential(names=["h"], types=[11], params=[(5e11, 15.0, 1.0, 1.0, 1.0)])
    qq = [3.0, -4.0, 1.5]
    assert np.array_equal(oracle.gradient(sph, qq), oracle.gradient(tri, qq)) and oracle.value(sph, qq) == oracle.value(tri, qq)
    gk_up, gk_dn = oracle.gradient(pots[5], [3.0, 1.0, 0.7]), oracle.gradient(pots[5], [3.0, 1.0, -0.7])
    assert gk_up[2] == -gk_dn[2] and gk_up[0] == gk_dn[0]


def test_leapfrog_reversible_and_energy(oracle):
    v2 = MilkyWayPotential("v2")
    w0 = np.array([8.0, 0.5, 0.1, 0.01, 0.22, 0.02])
    t = np.arange(0.0, 1001.0, 1.0)
    traj = oracle.leapfrog(v2, w0, t)
    E = 0.5 * (traj[3:] ** 2).sum(axis=0) + np.array([oracle.value(v2, traj[:3, j]) for j in range(traj.shape[1])])
    assert np.abs(E / E[0] - 1).max() < 1e-4          # symplectic: bounded energy error at dt = 1 Myr
    # reverse: flip the velocity of the synchronised end state and integrate the same grid
    wf = traj[:, -1].copy()
    wf[3:] *= -1
    back = oracle.leapfrog(v2, wf, t, store_all=False)
    back[3:] *= -1
    assert np.abs(back - w0).max() < 1e-10


def test_leapfrog_is_velocity_verlet(oracle):
    """The oracle's leapfrog (gala's recipe as recalled: half-step velocity initialised backwards, drift, gradient,
    kick, synchronised velocity by half a kick) against an INDEPENDENT textbook statement written here -- velocity
    Verlet / kick-drift-kick with the same gradient: the two are the same map up to rounding, so 1000 steps of a disc
    orbit agree to 1e-11 and every stored sample (not only the last) is the Verlet sample.  Pins the recipe's
    structure (what is updated with what, which velocity is stored); its rounding order stays [gala-recall]."""
    v2 = MilkyWayPotential("v2")
    w0 = np.array([8.0, 0.5, 0.1, 0.01, 0.22, 0.02])
    dt, n = 1.0, 1000
    traj = oracle.leapfrog(v2, w0, np.arange(0.0, (n + 1) * dt, dt))
    q, v = w0[:3].copy(), w0[3:].copy()
    a = -np.asarray(oracle.gradient(v2, q))
    worst = 0.0
    for j in range(1, n + 1):
        v_half = v + 0.5 * dt * a
        q = q + dt * v_half
        a = -np.asarray(oracle.gradient(v2, q))
        v = v_half + 0.5 * dt * a
        worst = max(worst, np.abs(traj[:3, j] - q).max() / np.abs(q).max(), np.abs(traj[3:, j] - v).max() / np.abs(v).max())
    assert worst < 1e-11, worst


# ---- dense-output DOP853 ("implement both modes") ---------------------------------

def test_dop853_dense_interpolant_matches_scipy(oracle):
    """Hairer's contd8 (oracle, dop853.c iout = 2) against scipy's independent DOP853 dense output of
    the SAME step: same tableau rows 14-16 and d4..d7, so the two polynomials agree to rounding."""
    from scipy.integrate import DOP853
    v2 = MilkyWayPotential("v2")

    def f(t, y):
        g = oracle.gradient(v2, y[:3])
        return np.concatenate([y[3:], -g])

    rng = np.random.default_rng(11)
    for trial in range(5):
        y0 = np.array([rng.uniform(2, 12), rng.uniform(-3, 3), rng.uniform(-1, 1),
                       rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.25), rng.uniform(-0.05, 0.05)])
        h = rng.uniform(0.5, 3.0)
        s = DOP853(f, 10.0, y0, 10.0 + 2 * h, first_step=h, rtol=1e-6, atol=1e-6)
        s.step()
        assert s.t == 10.0 + h            # one accepted step of h
        sol = s.dense_output()
        th = 10.0 + np.linspace(0.03, 0.97, 12) * h
        res, y, out, st = oracle.dop853_dense(v2, 10.0, y0, s.t, th, rtol=1e-6, atol=1e-6, h=h)
        assert res == 1 and st[0] == 1 and st[2] == 2 + 11 + 1 + 3     # 3 extra stages for the interpolant
        assert np.abs(y - s.y).max() < 1e-14
        assert np.abs(out - sol(th)).max() < 5e-15 * np.abs(out).max()


def test_dop853_dense_tableau_rows():
    """Rows 14-16 of the extended tableau sum to c14..c16 (0.1, 0.2, 7/9)."""
    import re
    txt = open(os.path.join(HERE, "..", "include", "cw_dop853_coeffs.h")).read()
    vals = {m.group(1): float(m.group(2)) for m in re.finditer(r"#define CW_(\w+) ([-\d.e+]+)", txt)}
    for i, c in ((14, 0.1), (15, 0.2), (16, 7.0 / 9.0)):
        row = sum(v for k, v in vals.items() if re.fullmatch(rf"A{i:02d}\d\d", k))
        assert abs(row - c) < 1e-14 and abs(vals[f"C{i}"] - c) < 1e-15
    assert sum(1 for k in vals if re.fullmatch(r"D[4-7]\d\d", k)) == 48


def test_dop853_dense_vs_restart_and_kat5(oracle):
    """The two DOPRI853 modes solve the same problem to the same tolerance: trajectories agree to
    ~1e-8 relative over 0.5 Gyr at rtol = atol = 1e-10 (north_star allows 1e-7), the dense mode needs
    several times fewer gradient evaluations, and it also reproduces the notebook's printed samples."""
    v2 = MilkyWayPotential("v2")
    y0 = np.array([8.0, 0.3, 0.1, 0.01, 0.22, 0.02])
    r1 = oracle.integrate_orbit(v2, y0, 0.0, 500.0, 1.0, integrator=oracle.DOP853, store_all=True)
    r2 = oracle.integrate_orbit(v2, y0, 0.0, 500.0, 1.0, integrator=oracle.DOP853_DENSE, store_all=True)
    assert r1["status"] == 0 and r2["status"] == 0 and r1["n_nodes"] == r2["n_nodes"] == 500
    assert np.array_equal(r1["t"], r2["t"])
    # relative to the orbit's own size (|pos|, |vel|), as in the GPU parity tests
    sp, sv = np.linalg.norm(r1["traj"][:3], axis=0), np.linalg.norm(r1["traj"][3:], axis=0)
    d = max((np.abs(r1["traj"][:3] - r2["traj"][:3]).max(axis=0) / sp).max(),
            (np.abs(r1["traj"][3:] - r2["traj"][3:]).max(axis=0) / sv).max())
    assert d < 1e-7, d
    assert np.array_equal(r2["traj"][:, -1], r2["w_final"]) and np.array_equal(r2["traj"][:, 0], y0)
    assert r2["stats"][2] < 0.5 * r1["stats"][2]
    # final state only == last sample of the stored trajectory, bit for bit (the interpolant never feeds back)
    r3 = oracle.integrate_orbit(v2, y0, 0.0, 500.0, 1.0, integrator=oracle.DOP853_DENSE, store_all=False)
    assert np.array_equal(r3["w_final"], r2["w_final"])
    # KAT-5 through the dense path
    import ctypes as C
    P = oracle.make_potential(v2)
    dp = C.POINTER(C.c_double)
    for pos, vel, t in ((KATS["KAT5_pos_first3_kpc"], KATS["KAT5_vel_first3_kpc_per_myr"], KATS["KAT4_t_first3_myr"]),
                        (KATS["KAT5_pos_last3_kpc"], KATS["KAT5_vel_last3_kpc_per_myr"], KATS["KAT4_t_last3_myr"])):
        w = np.array(pos[0] + vel[0])
        tt = np.array(t)
        out = np.empty((6, 3))
        rc = oracle.lib().cwo_dop853_dense_segment(C.byref(P), w.ctypes.data_as(dp), tt.ctypes.data_as(dp), 3, 1e-10,
                                                   1e-10, 0, out.ctypes.data_as(dp), 3, None)
        assert rc == 0
        exp = np.array([pos[1] + vel[1], pos[2] + vel[2]])
        assert np.abs(out.T[1:] - exp).max() < 2.5e-8


def test_dop853_dense_driver_with_kicks(oracle):
    """Kick nodes, node counts and the stitched trajectory layout do not depend on the DOPRI853 mode."""
    from cogsworth_b200.synthetic import make_population
    pop = make_population(40, cfg_id=7, horizon_gyr=0.15)
    b = pop.batch
    kw = dict(ev_off=b.ev_off, ev_time=b.ev_time, ev_dv=b.ev_dv)
    r0 = oracle.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, integrator=oracle.DOP853, **kw)
    off = np.concatenate([[0], np.cumsum(r0["n_nodes"])]).astype(np.int64)
    r1 = oracle.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, integrator=oracle.DOP853,
                                traj_off=off, **kw)
    r2 = oracle.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, integrator=oracle.DOP853_DENSE,
                                traj_off=off, **kw)
    assert (r1["status"] == 0).all() and (r2["status"] == 0).all()
    assert np.array_equal(r1["ev_step"], r2["ev_step"]) and np.array_equal(r1["n_nodes"], r2["n_nodes"])
    assert np.array_equal(r1["traj_t"], r2["traj_t"]) and not np.isnan(r2["traj_pos"]).any()
    scale = np.abs(r1["traj_pos"]).max()
    assert np.abs(r1["traj_pos"] - r2["traj_pos"]).max() / scale < 1e-7
    assert np.array_equal(r2["traj_pos"][:, off[1:] - 1], r2["w_final"][:3])
    # the sample stored on a kick node carries the kicked velocity in both modes
    for i in np.flatnonzero(b.ev_off[1:] > b.ev_off[:-1])[:10]:
        e = b.ev_off[i]
        k = r2["ev_step"][e]
        if k == 0:
            continue
        j1 = r1["traj_vel"][:, off[i] + k] - r1["traj_vel"][:, off[i] + k - 1]
        j2 = r2["traj_vel"][:, off[i] + k] - r2["traj_vel"][:, off[i] + k - 1]
        assert np.abs(j1 - j2).max() < 1e-8 and np.abs(j2).max() > 0
