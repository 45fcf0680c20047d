"""The general kernels (POT_GENERAL: full gradient vector, per-component rotation, time-interpolated masses) against
the CPU oracle: the rest of the web UI's potential menu (reference online-interface/potentials.py:93-131: LM10,
BovyMWPotential2014, LongMuraliBar, LeeSutoTriaxialNFW) and gp.TimeInterpolatedPotential
(docs/tutorials/potentials/evolving_potentials.ipynb cells 5, 13), SURVEY 8(f) N4.
Tolerances as everywhere: leapfrog 1e-10, DOPRI853 1e-7 relative, kick nodes bit-exact.
"""
import numpy as np
import pytest

from cogsworth_b200 import _lib
from cogsworth_b200 import potential as P
from cogsworth_b200.synthetic import make_population

from test_gpu_parity import check_bit_exact_bookkeeping, rel_diff, run_both, _local_step_check

pytestmark = pytest.mark.gpu


def _population(n, cfg_id, potential, horizon_gyr=0.3):
    # birth velocities are drawn in the static Milky Way, the orbits integrated in `potential`
    pop = make_population(n, cfg_id=cfg_id, potential=P.MilkyWayPotential("v2"), horizon_gyr=horizon_gyr)
    pop.potential = potential
    return pop


def _menu():
    bar = P.CompositePotential(
        names=["disk", "bulge", "bar", "halo"], types=[1, 2, P.COMP_LONG_MURALI_BAR, 3],
        params=[(5e10, 3.0, 0.28), (4e9, 1.0, 0.0), (1e10, 4.0, 1.0, 0.5), (5.4e11, 15.62, 0.0)],
        rotations=[None, None, P.rotation_z(np.deg2rad(25.0)), None], label="barred")
    leesuto = P.CompositePotential(
        names=["disk", "halo"], types=[1, P.COMP_LEESUTO_NFW], params=[(6e10, 3.0, 0.28), (0.2, 20.0, 1.0, 0.9, 0.8)],
        label="leesuto")
    return {"lm10": P.LM10Potential(), "bovy2014": P.BovyMWPotential2014(), "barred": bar, "leesuto": leesuto}


@pytest.mark.parametrize("name", ["lm10", "bovy2014", "barred", "leesuto"])
def test_menu_potentials_pointwise_and_integrated(ctx, oracle, name):
    pot = _menu()[name]
    ctx.set_potential(pot)
    rng = np.random.default_rng(21)
    q = rng.normal(0, 7, (3, 1024))
    g, phi, vc = ctx.gradient(q), ctx.potential_value(q), ctx.circular_velocity(q)
    for i in range(0, 1024, 16):
        go = oracle.gradient(pot, q[:, i].copy())
        assert np.allclose(g[:, i], go, rtol=2e-12, atol=1e-15 * np.abs(go).max()), (name, i, g[:, i], go)
        assert abs(phi[i] - oracle.value(pot, q[:, i].copy())) < 2e-12 * abs(phi[i])
        assert abs(vc[i] - oracle.circular_velocity(pot, q[:, i].copy())) < 2e-12 * vc[i]
    pop = _population(200, 31, pot)
    for integ, tol in ((_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7), (_lib.DOP853_DENSE, 1e-7)):
        g_, o_ = run_both(ctx, oracle, pop, integ)
        check_bit_exact_bookkeeping(g_, o_)
        assert (g_["status"] == 0).all()
        d = rel_diff(g_["w_final"], o_["w_final"])
        assert np.quantile(d, 0.9) < tol, (name, integ, np.quantile(d, [0.5, 0.9, 1]))
    # Lee & Suto's F2, F3 cancel like 1/u^4 at small u = r / r_s (in gala's formula as here), so two correct
    # evaluations of a step through the centre differ by ~1e-7 in velocity
    _local_step_check(ctx, oracle, pop, name, tol_v=1e-6 if name == "leesuto" else 1e-9)


def test_rotated_and_lm10_kinds_return_the_general_kernels_bits(ctx, monkeypatch):
    """LM10 (static, diagonal components, one rotated frame) runs in kernels compiled for its component list (POT_LM10);
    other such composites in the lighter POT_ROTATED kernels.  Both must return what the POT_GENERAL kernels return for
    the same potential, bit for bit, in every integrator and pointwise."""
    pot = _menu()["lm10"]
    pop = _population(300, 77, pot)
    q = np.random.default_rng(5).normal(0, 9, (3, 512))
    out = {}
    for kind, env in (("lm10", None), ("rotated", "CW_NO_LM10_KIND"), ("general", "CW_NO_ROTATED_KIND")):
        if env:
            monkeypatch.setenv(env, "1")     # read when the potential is set
        ctx.set_potential(pot)
        res = [ctx.gradient(q), ctx.potential_value(q)]
        for integ in (_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE):
            r = ctx.integrate_batch(pop.batch, integrator=integ, atol=1e-10, rtol=1e-10)
            res += [r["w_final"], r["status"], r["stats"]]
        r = ctx.integrate_batch(pop.batch, integrator=_lib.LEAPFROG, store_all=True)
        res += [r["traj_pos"][:, : r["traj_off"][-1]], r["traj_vel"][:, : r["traj_off"][-1]]]
        out[kind] = res
    monkeypatch.delenv("CW_NO_LM10_KIND")
    monkeypatch.delenv("CW_NO_ROTATED_KIND")
    for kind in ("lm10", "rotated"):
        for k, (a, b) in enumerate(zip(out[kind], out["general"])):
            assert np.array_equal(a, b), (kind, k)


def test_rotation_is_a_change_of_frame(ctx):
    """An orbit in a rotated bar = the rotated orbit in the unrotated bar (and R = identity changes nothing)."""
    ang = 0.7
    R = P.rotation_z(ang)
    bar0 = P.CompositePotential(names=["bar", "halo"], types=[P.COMP_LONG_MURALI_BAR, 3],
                                params=[(2e10, 5.0, 1.5, 0.6), (5e11, 16.0, 0.0)])
    barR = P.CompositePotential(names=bar0.names, types=bar0.types, params=bar0.params, rotations=[R, None])
    barI = P.CompositePotential(names=bar0.names, types=bar0.types, params=bar0.params, rotations=[np.eye(3), None])
    rng = np.random.default_rng(4)
    n = 256
    w = np.vstack([rng.normal(0, 5, (3, n)), rng.normal(0, 0.12, (3, n))])
    wr = np.vstack([R @ w[:3], R @ w[3:]])          # the same orbits seen from the bar's frame
    ctx.set_potential(barR)
    a = ctx.integrate(w, 0.0, 300.0, 1.0, 1.0, 0, integrator=_lib.LEAPFROG)
    ctx.set_potential(bar0)
    b = ctx.integrate(wr, 0.0, 300.0, 1.0, 1.0, 0, integrator=_lib.LEAPFROG)
    back = np.vstack([R.T @ b["w_final"][:3], R.T @ b["w_final"][3:]])
    d = rel_diff(a["w_final"], back)
    assert np.quantile(d, 0.9) < 1e-10, np.quantile(d, [0.5, 0.9, 1])
    ctx.set_potential(barI)
    c = ctx.integrate(wr, 0.0, 300.0, 1.0, 1.0, 0, integrator=_lib.LEAPFROG)
    assert np.array_equal(c["w_final"], b["w_final"])


def _evolving(interp, shape):
    """shape = "common": one mass-fraction curve for every component (the tutorial; separable, Phi = f(t) Phi_0: the
    Milky Way kernel with a scale factor).  "uneven knots": the same on unequally spaced knots (binary search).
    "per-component": the halo grows on its own curve (the general kernels, one spline per component)."""
    f = np.array([0.1, 0.2, 0.55, 0.8, 1.0])
    if shape == "common":
        return P.evolving_milky_way(np.linspace(0, 12000, 5), f, interp=interp)
    if shape == "uneven knots":
        return P.evolving_milky_way(np.array([0.0, 2000.0, 5000.0, 9500.0, 12000.0]), f, interp=interp)
    base = P.evolving_milky_way([0.0, 12000.0], [1.0, 1.0])
    base = P.CompositePotential(names=base.names, types=base.types, params=base.params)
    g = np.array([0.3, 0.35, 0.5, 0.9, 1.0])
    amps = [p[0] * (g if i == 5 else f) for i, p in enumerate(base.params)]
    return P.time_interpolated(base, np.linspace(0, 12000, 5), amps, interp=interp)


@pytest.mark.parametrize("interp,shape", [("cubic", "common"), ("linear", "common"), ("cubic", "uneven knots"),
                                          ("cubic", "per-component"), ("linear", "per-component")])
def test_evolving_milky_way(ctx, oracle, interp, shape):
    """evolving_potentials.ipynb cell 13: every component's mass grows from 10 % to 100 % over 12 Gyr."""
    pot = _evolving(interp, shape)
    ctx.set_potential(pot.at_time(7000.0))
    q = np.random.default_rng(2).normal(0, 7, (3, 256))
    g = ctx.gradient(q)
    for i in range(0, 256, 8):
        assert np.allclose(g[:, i], oracle.gradient_t(pot, 7000.0, q[:, i].copy()), rtol=1e-12)
    # Wagg2022 birth times: every orbit runs on its own stretch of the 12 Gyr of mass growth
    pop = _population(150, 33, pot, horizon_gyr=None)
    for integ, tol in ((_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7), (_lib.DOP853_DENSE, 1e-7)):
        g_, o_ = run_both(ctx, oracle, pop, integ)
        check_bit_exact_bookkeeping(g_, o_)
        d = rel_diff(g_["w_final"], o_["w_final"])
        # linear interpolation puts kinks into m(t): the dense form takes steps of many Myr across them and its
        # accept / reject decisions there amplify last-bit differences (median stays ~1e-9)
        kinked = interp == "linear" and integ == _lib.DOP853_DENSE
        assert np.quantile(d, 0.5 if kinked else 0.8) < tol, (interp, shape, integ, np.quantile(d, [0.5, 0.8, 0.9, 1]))
    # the evolution is not a detail: the same orbits in the static potential end somewhere else
    static = _population(150, 33, P.evolving_milky_way([0.0, 12000.0], [1.0, 1.0]), horizon_gyr=None)
    s_, _ = run_both(ctx, oracle, static, _lib.LEAPFROG)
    assert np.median(np.linalg.norm(s_["w_final"][:3] - g_["w_final"][:3], axis=0)) > 0.5
    _local_step_check(ctx, oracle, _population(60, 34, pot, horizon_gyr=0.4), "evolving MW")


def test_constant_knots_reproduce_the_static_kernels(ctx):
    """MilkyWayPotential2022 with every mass 'interpolated' through constant knots runs in the general kernels and
    must agree with the specialised Milky Way kernels (different arithmetic, same orbits)."""
    mw = P.MilkyWayPotential("v2")
    const = P.time_interpolated(mw, [0.0, 6000.0, 12000.0], [np.full(3, p[0]) for p in mw.params])
    pop = make_population(400, cfg_id=35, potential=mw, horizon_gyr=0.3)
    b = pop.batch
    out = {}
    for key, pot in (("mw", mw), ("const", const)):
        ctx.set_potential(pot)
        out[key] = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                                 integrator=_lib.LEAPFROG)
    assert np.array_equal(out["mw"]["ev_step"], out["const"]["ev_step"])
    d = rel_diff(out["const"]["w_final"], out["mw"]["w_final"])
    assert np.quantile(d, 0.9) < 1e-10, np.quantile(d, [0.5, 0.9, 1])


def test_general_entry_point_validates(ctx):
    bad_R = P.CompositePotential(names=["h"], types=[3], params=[(1e12, 10.0, 0.0)],
                                 rotations=[np.array([[1.0, 0.2, 0], [0, 1.0, 0], [0, 0, 1.0]])])
    with pytest.raises(RuntimeError, match="orthogonal"):
        ctx.set_potential(bad_R)
    base = P.CompositePotential(names=["h"], types=[3], params=[(1e12, 10.0, 0.0)])
    with pytest.raises(RuntimeError, match="increasing"):
        ctx.set_potential(P.time_interpolated(base, [0.0, 5.0, 5.0], [np.ones(3)]))
    with pytest.raises(NotImplementedError):
        ctx.set_potential(P.time_interpolated(base, np.arange(100.0), [np.ones(100)]))
    log = P.CompositePotential(names=["l"], types=[8], params=[(0.2, 10.0, 1.0)])
    with pytest.raises(RuntimeError, match="v_c"):
        ctx.set_potential(P.time_interpolated(log, [0.0, 1.0], [np.array([0.1, 0.2])]))
    ctx.set_potential(P.MilkyWayPotential("v2"))


def test_tutorial_evolving_nfw_through_the_reference_signature(ctx, oracle):
    """evolving_potentials.ipynb cells 4-8: one test particle, 12 Gyr at dt = 1 Myr with the default DOPRI853, in the
    static and in the growing NFW halo -- through integrate_orbit_with_events, checked against the oracle and against
    scipy's DOP853 of the same time-dependent equations of motion."""
    import cogsworth_b200 as cb
    from scipy.integrate import solve_ivp
    static = P.CompositePotential(names=["halo"], types=[3], params=[(1e12, 10.0, 0.0)], label="NFW")
    evolving = P.time_interpolated(static, np.linspace(0, 12000, 20), [np.linspace(1e11, 1e12, 20)])
    w0 = [8.0, 0.0, 0.0, 50.0, 50.0, -5.0]                       # kpc, km/s (cell 7)
    k = 1.0 / 977.7922216807891
    w_gal = np.array(w0[:3] + [v * k for v in w0[3:]])
    ends = {}
    for name, pot in (("static", static), ("evolving", evolving)):
        orb = cb.integrate_orbit_with_events(w0, 0.0, 12.0, 1.0, potential=pot, events=None, store_all=True, context=ctx)
        assert orb.shape == (12000,) and abs(orb.t[-1] - 11999.0) < 1e-9
        ref = oracle.integrate_orbit(pot, w_gal, 0.0, 12000.0, 1.0, integrator=oracle.DOP853, store_all=True)
        assert np.abs(orb.pos.xyz - ref["traj"][:3]).max() < 1e-6
        sol = solve_ivp(lambda t, y: np.concatenate([y[3:], -oracle.gradient_t(pot, t, y[:3])]), (0.0, 11999.0), w_gal,
                        method="DOP853", rtol=1e-13, atol=1e-15)
        assert np.abs(orb.pos.xyz[:, -1] - sol.y[:3, -1]).max() < 1e-5      # 12 000 restarted intervals at the default 1e-10
        ends[name] = orb.pos.xyz[:, -1]
    assert np.linalg.norm(ends["static"] - ends["evolving"]) > 1.0       # "the orbits trace very different paths"


@pytest.mark.parametrize("k", [0, 1, 2])
def test_cuda_path_reproduces_the_frozen_general_cases(ctx, k):
    """The committed fixtures of tests/golden/oracle_cases.npz (tools/gen_golden_oracle.py) for LM10, BovyMWPotential2014
    and the per-component evolving Milky Way: kick nodes bit-exact, final states within the north_star tolerances."""
    import os
    from _cases import general_cases
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_cases.npz"))
    name, pot, integ, horizon = general_cases()[k]
    b = make_population(48, cfg_id=1, potential=P.MilkyWayPotential("v2"), horizon_gyr=horizon).batch
    ctx.set_potential(pot)
    r = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["ev_step"], G[name + "_ev_step"]) and np.array_equal(r["n_nodes"], G[name + "_n_nodes"])
    d = rel_diff(r["w_final"], G[name + "_w_final"])
    assert np.quantile(d, 0.9) < (1e-10 if integ == _lib.LEAPFROG else 1e-7), (name, np.quantile(d, [0.5, 0.9, 1]))
