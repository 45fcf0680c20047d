"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs, at
sizes the oracle finishes in seconds; plus the notebook known-answers and the frozen golden cases.

Tolerances (BASELINE.json north_star): leapfrog <= 1e-10 relative in final phase-space coordinates,
DOPRI853 <= 1e-7 relative, kick node index bit-exact.  The kernels use a different but equivalent
arithmetic (fused rsqrt/reciprocal/log, FMA contraction), and leapfrog at dt = 1 Myr does not resolve
nucleus-grazing orbits (Hernquist c = 0.07 kpc): those amplify last-bit differences chaotically
(SURVEY.md 7 "hard parts": 0.5 % of orbits beyond 1e-10 after 1000 steps between ANY two roundings).
So the leapfrog assertions are made on the orbits that do NOT amplify last-bit noise -- classified by the oracle
itself, re-run with rounding noise in its gradient (tests/_cases.py: sensitive_orbits): 100 % of those must meet the
tolerance, i.e. every orbit beyond it is one the oracle cannot reproduce against its own rounding.
"""
import json
import os

import numpy as np
import pytest

import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200 import constants as K
from cogsworth_b200.potential import CompositePotential, MilkyWayPotential
from cogsworth_b200.synthetic import make_population

from _cases import assert_parity_outside_the_sensitive_set, rel_diff, sensitive_orbits

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_both(ctx, O, pop, integ, **kw):
    b = pop.batch
    ctx.set_potential(pop.potential)
    g = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ, **kw)
    okw = {k: v for k, v in kw.items() if k in ("atol", "rtol", "nmax")}
    o = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                          integrator=integ, traj_off=g.get("traj_off") if kw.get("store_all") else None, **okw)
    return g, o


def min_radius(O, pop, idx):
    """Pericentre of selected orbits from the oracle's stored trajectory."""
    b = pop.batch.take(idx)
    nn = np.array([O.lib().cwo_time_grid(b.t1[i], b.t2[i], b.dt[i], int(b.flags[i]) & 1, None, 0) for i in range(len(idx))])
    off = np.concatenate([[0], np.cumsum(nn)])
    o = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                          integrator=O.LEAPFROG, traj_off=off)
    r = np.linalg.norm(o["traj_pos"], axis=0)
    return np.array([r[off[i]:off[i + 1]].min() for i in range(len(idx))])


def check_bit_exact_bookkeeping(g, o):
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["n_nodes"], o["n_nodes"])
    assert np.array_equal(g["ev_step"], o["ev_step"]), "kick node indices must be bit-exact"


def test_config1_leapfrog_v1(ctx, oracle):
    """configs[0]: 1000 binaries, MilkyWayPotential v1, 200 Myr leapfrog dt = 1 Myr with kicks."""
    pop = make_population(1000, cfg_id=1, potential=MilkyWayPotential("v1"), horizon_gyr=0.2)
    g, o = run_both(ctx, oracle, pop, _lib.LEAPFROG)
    check_bit_exact_bookkeeping(g, o)
    assert (g["status"] == 0).all() and (g["n_nodes"] == 201).all()
    d = rel_diff(g["w_final"], o["w_final"])
    sens, _, _ = sensitive_orbits(oracle, pop.potential, pop.batch, oracle.LEAPFROG, ref=o)
    rep = assert_parity_outside_the_sensitive_set(d, sens, 1e-10, 0.03, "config 1")
    assert rep["median"] < 1e-13, rep
    out = np.flatnonzero(d >= 1e-10)
    if len(out):
        assert (min_radius(oracle, pop, out) < 1.0).all(), "only nucleus-grazing orbits may exceed 1e-10"
    assert ctx.last_launch_count <= 3          # (unpack) + grid pre-pass + integrator
    # frozen golden case (same generator, 96 binaries)
    G = np.load(os.path.join(HERE, "golden", "oracle_cases.npz"))
    pop96 = make_population(96, cfg_id=1, potential=MilkyWayPotential("v1"), horizon_gyr=0.2)
    g96, _ = run_both(ctx, oracle, pop96, _lib.LEAPFROG)
    assert np.array_equal(g96["ev_step"], G["cfg1_leapfrog_v1_ev_step"])
    d96 = rel_diff(g96["w_final"], G["cfg1_leapfrog_v1_w_final"])
    s96, _, _ = sensitive_orbits(oracle, pop96.potential, pop96.batch, oracle.LEAPFROG)
    assert_parity_outside_the_sensitive_set(d96, s96, 1e-10, 0.05, "config 1 golden")
    # the frozen DOPRI853 cases (MW v2, same generator), both forms
    pop96v2 = make_population(96, cfg_id=1, potential=MilkyWayPotential("v2"), horizon_gyr=0.2)
    for name, integ in (("cfg1_dop853_v2", _lib.DOP853), ("cfg1_dop853dense_v2", _lib.DOP853_DENSE)):
        gd, _ = run_both(ctx, oracle, pop96v2, integ)
        assert np.array_equal(gd["ev_step"], G[name + "_ev_step"]) and np.array_equal(gd["n_nodes"], G[name + "_n_nodes"])
        dd = rel_diff(gd["w_final"], G[name + "_w_final"])
        assert np.quantile(dd, 0.9) < 1e-7, (name, np.quantile(dd, [0.5, 0.9, 1]))


@pytest.mark.parametrize("integ,tol", [(_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7)])
def test_config2_wagg_horizons_v2(ctx, oracle, integ, tol):
    """configs[1] scaled to 250 binaries: MilkyWayPotential2022, per-system Wagg2022 horizons
    (up to 12 000 steps, divergent step counts), kicks, final positions only."""
    pop = make_population(250, cfg_id=2)
    g, o = run_both(ctx, oracle, pop, integ)
    check_bit_exact_bookkeeping(g, o)
    assert g["n_nodes"].max() > 9000 and g["n_nodes"].min() < 3000
    d = rel_diff(g["w_final"], o["w_final"])
    # long horizons: the share of orbits that amplify rounding noise grows with the step count (measured with the
    # oracle's own noise: 5 % beyond 1e-11 after 1000 steps, 22-25 % after the Wagg2022 horizons of up to 12 000)
    assert np.median(d) < (1e-11 if integ == _lib.LEAPFROG else 1e-9), np.quantile(d, [0.5, 0.8, 0.9, 1])
    if integ == _lib.LEAPFROG:
        sens, _, _ = sensitive_orbits(oracle, pop.potential, pop.batch, oracle.LEAPFROG, ref=o)
        assert_parity_outside_the_sensitive_set(d, sens, tol, 0.35, "config 2 leapfrog")
    else:
        assert np.quantile(d, 0.8) < tol, np.quantile(d, [0.5, 0.8, 0.9, 1])
    if integ == _lib.DOP853:
        # same controller: accepted / rejected step counts agree to a fraction of a percent
        assert abs(g["stats"][0] - o["stats"][0]) <= 2e-3 * o["stats"][0]
        assert g["stats"][3] == int((g["n_nodes"] - 1).sum())


def test_config3_store_all_trajectories(ctx, oracle):
    """configs[2] scaled: full trajectories in the orbits/{offsets,pos,vel,t} layout."""
    pop = make_population(400, cfg_id=3, horizon_gyr=0.25)
    g, o = run_both(ctx, oracle, pop, _lib.LEAPFROG, store_all=True)
    check_bit_exact_bookkeeping(g, o)
    off = g["traj_off"]
    assert off[-1] == int(g["n_nodes"].sum()) and not np.isnan(g["traj_pos"]).any()
    assert np.array_equal(g["traj_t"], o["traj_t"])           # time stamps are bit-exact
    # last sample of each orbit is w_final
    assert np.array_equal(g["traj_pos"][:, off[1:] - 1], g["w_final"][:3])
    assert np.array_equal(g["traj_vel"][:, off[1:] - 1], g["w_final"][3:])
    # sample 0 is the (possibly kicked) initial state, bit-exact
    first = off[:-1]
    assert np.array_equal(g["traj_pos"][:, first], pop.batch.w0[:3])
    dp = np.abs(g["traj_pos"] - o["traj_pos"]).max(axis=0)
    assert np.median(dp) < 1e-12 and np.quantile(dp, 0.9) < 1e-9
    # no duplicated stitch samples (reference tests/test_kicks.py:7-38)
    for i in np.flatnonzero(pop.batch.ev_off[1:] > pop.batch.ev_off[:-1])[:50]:
        seg = slice(off[i], off[i + 1])
        assert (np.diff(g["traj_t"][seg]) != 0).all() and (np.diff(g["traj_pos"][0, seg]) != 0).all()
    # disrupted secondaries share the primary's trajectory bit for bit up to the disruption node (plot.py:558)
    b = pop.batch
    n = b.n_primary
    checked = 0
    for s, prim in enumerate(b.secondary_of[:60]):
        sec = n + s
        ks = g["ev_step"][b.ev_off[sec]:b.ev_off[sec + 1]]
        k = ks[-1]
        assert np.array_equal(g["traj_pos"][:, off[sec]:off[sec] + k], g["traj_pos"][:, off[prim]:off[prim] + k])
        assert np.array_equal(g["traj_vel"][:, off[sec]:off[sec] + k], g["traj_vel"][:, off[prim]:off[prim] + k])
        checked += 1
    assert checked > 10
    # DOPRI853 trajectories too
    g2, o2 = run_both(ctx, oracle, pop, _lib.DOP853, store_all=True)
    assert np.array_equal(g2["traj_t"], o2["traj_t"])
    dp2 = np.abs(g2["traj_pos"] - o2["traj_pos"]).max(axis=0)
    assert np.quantile(dp2, 0.95) < 1e-7


def test_config5_dop853_plunging(ctx, oracle):
    """configs[4] scaled: DOPRI853 rtol = atol = 1e-8, disrupted-secondary kicks, 5 % plunging orbits."""
    pop = make_population(300, cfg_id=5, horizon_gyr=1.0, plunging_fraction=0.05)
    g, o = run_both(ctx, oracle, pop, _lib.DOP853, atol=1e-8, rtol=1e-8)
    check_bit_exact_bookkeeping(g, o)
    assert pop.batch.n_orbits - pop.batch.n_primary > 0.4 * pop.batch.n_primary
    d = rel_diff(g["w_final"], o["w_final"])
    assert np.median(d) < 1e-9 and np.quantile(d, 0.9) < 1e-7, np.quantile(d, [0.5, 0.9, 0.99, 1])
    acc_g, rej_g, nf_g, ni_g = [int(x) for x in g["stats"]]
    acc_o, rej_o, nf_o = o["stats"]
    assert ni_g == int((g["n_nodes"] - 1).sum())
    assert abs(acc_g - acc_o) <= 5e-3 * acc_o and abs(rej_g - rej_o) <= max(20, 0.05 * rej_o)
    assert acc_g > ni_g          # sub-stepping happened (divergent step counts)


@pytest.mark.parametrize("integ", [_lib.LEAPFROG, _lib.DOP853])
def test_results_independent_of_queue_order_and_chunking(ctx, oracle, integ):
    """Large batches are reordered (longest / most expensive orbit first; DOP853 runs a pilot to
    estimate the cost) and host-buffer calls are cut into pipelined chunks: neither may change a bit
    of any orbit's result."""
    pop = make_population(3400, cfg_id=41, horizon_gyr=0.12, plunging_fraction=0.05)
    b = pop.batch
    assert b.n_orbits > 4096                       # above the sort threshold
    ctx.set_potential(pop.potential)
    big = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ)
    idx = np.arange(500, 1500)
    sub = b.take(idx)
    small = ctx.integrate(sub.w0, sub.t1, sub.t2, sub.dt, sub.to_myr, sub.flags, sub.ev_off, sub.ev_time, sub.ev_dv,
                          integrator=integ)
    assert np.array_equal(big["w_final"][:, idx], small["w_final"])
    assert np.array_equal(big["status"][idx], small["status"])
    o = oracle.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                               integrator=integ)
    check_bit_exact_bookkeeping(big, o)
    d = rel_diff(big["w_final"], o["w_final"])
    if integ == _lib.LEAPFROG:
        sens, _, _ = sensitive_orbits(oracle, pop.potential, b, oracle.LEAPFROG, ref=o)
        assert_parity_outside_the_sensitive_set(d, sens, 1e-10, 0.08, "queue order / chunking")
    else:
        assert np.quantile(d, 0.9) < 1e-7
    if integ == _lib.DOP853:
        assert abs(int(big["stats"][0]) - o["stats"][0]) <= 5e-3 * o["stats"][0]


def test_generic_composite_kernel(ctx, oracle):
    """A component list that is neither MW shape runs through the generic kernel."""
    pot = CompositePotential(names=["bulge", "disk", "halo", "disk2"], types=[2, 1, 3, 1],
                             params=[(4e9, 0.8, 0.0), (5e10, 3.2, 0.25), (6e11, 17.0, 0.0), (1e10, 5.0, 0.5)])
    pop = make_population(300, cfg_id=11, potential=pot, horizon_gyr=0.3)
    for integ, tol in ((_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7)):
        g, o = run_both(ctx, oracle, pop, integ)
        check_bit_exact_bookkeeping(g, o)
        d = rel_diff(g["w_final"], o["w_final"])
        assert np.quantile(d, 0.95) < tol, np.quantile(d, [0.5, 0.95, 1])


def test_generic_kernel_extra_components(ctx, oracle):
    """Kepler, Plummer, Isochrone, Jaffe and the axisymmetric Logarithmic halo through the generic kernel
    (gradient, value, v_circ pointwise; leapfrog and both DOPRI853 forms with kicks)."""
    pot = CompositePotential(names=["disk", "bh", "bulge", "iso", "jaffe", "halo", "kuzmin", "burkert"],
                             types=[1, 4, 5, 6, 7, 8, 9, 10],
                             params=[(5e10, 3.0, 0.28), (4e6, 0.0, 0.0), (6e9, 0.6, 0.0), (8e9, 1.5, 0.0),
                                     (2e10, 4.0, 0.0), (0.19, 8.0, 0.85), (1e10, 3.5, 0.0), (1.5e7, 9.0, 0.0)])
    ctx.set_potential(pot)
    rng = np.random.default_rng(12)
    q = rng.normal(0, 7, (3, 2048))
    g = ctx.gradient(q)
    phi = ctx.potential_value(q)
    vc = ctx.circular_velocity(q)
    for i in range(0, 2048, 16):
        go = oracle.gradient(pot, q[:, i].copy())
        assert np.allclose(g[:, i], go, rtol=1e-13, atol=1e-18)
        assert abs(phi[i] - oracle.value(pot, q[:, i].copy())) < 1e-13 * abs(phi[i])
        assert abs(vc[i] - oracle.circular_velocity(pot, q[:, i].copy())) < 1e-13 * vc[i]
    # the razor-thin Kuzmin disk has a kink in g_z at z = 0: fine for the fixed-step leapfrog, but an adaptive
    # integrator's accept/reject decisions at the crossing amplify last-bit differences, so the DOPRI853 forms
    # are compared in the same composite without it
    smooth = CompositePotential(names=pot.names[:6] + pot.names[7:], types=pot.types[:6] + pot.types[7:],
                                params=pot.params[:6] + pot.params[7:])
    for integ, tol in ((_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7), (_lib.DOP853_DENSE, 1e-7)):
        pop = make_population(200, cfg_id=13, potential=pot if integ == _lib.LEAPFROG else smooth, horizon_gyr=0.25)
        g_, o_ = run_both(ctx, oracle, pop, integ)
        check_bit_exact_bookkeeping(g_, o_)
        d = rel_diff(g_["w_final"], o_["w_final"])
        assert np.quantile(d, 0.9) < tol, (integ, np.quantile(d, [0.5, 0.9, 1]))


def test_flattened_nfw_from_the_reference_tutorial(ctx, oracle):
    """docs/tutorials/pop_settings/potentials.ipynb:141 swaps the Milky Way for gp.NFWPotential(m=6.09e11, r_s=10,
    a=0.25): a halo flattened along x, i.e. grad = (Fx x, Fy y, Fz z) with Fx != Fy.  Pointwise and integrated
    parity against the oracle, alone and inside a composite."""
    tut = CompositePotential(names=["nfw"], types=[11], params=[(6.09e11, 10.0, 0.25, 1.0, 1.0)])
    mix = CompositePotential(names=["disk", "bulge", "halo", "loghalo"], types=[1, 2, 11, 12],
                             params=[(6e10, 3.0, 0.28), (5e9, 1.0, 0.0), (3e11, 16.0, 1.0, 0.85, 0.7),
                                     (0.12, 12.0, 1.38, 1.0, 1.36)])
    rng = np.random.default_rng(3)
    q = rng.normal(0, 7, (3, 1024))
    for pot in (tut, mix):
        ctx.set_potential(pot)
        g, phi = ctx.gradient(q), ctx.potential_value(q)
        for i in range(0, 1024, 16):
            assert np.allclose(g[:, i], oracle.gradient(pot, q[:, i].copy()), rtol=1e-13, atol=1e-18)
            assert abs(phi[i] - oracle.value(pot, q[:, i].copy())) < 1e-13 * abs(phi[i])
        pop = make_population(200, cfg_id=15, potential=pot, horizon_gyr=0.3)
        for integ, tol in ((_lib.LEAPFROG, 1e-10), (_lib.DOP853, 1e-7), (_lib.DOP853_DENSE, 1e-7)):
            g_, o_ = run_both(ctx, oracle, pop, integ)
            check_bit_exact_bookkeeping(g_, o_)
            d = rel_diff(g_["w_final"], o_["w_final"])
            assert np.quantile(d, 0.9) < tol, (pot.names, integ, np.quantile(d, [0.5, 0.9, 1]))
        _local_step_check(ctx, oracle, pop, "flattened NFW")
    # the three-parameter entry point still refuses it (it cannot carry a, b, c)
    with pytest.raises(RuntimeError):
        bad = CompositePotential(names=["nfw"], types=[11], params=[(6.09e11, 10.0, 0.25)])
        bad.param_array = lambda: np.array([[6.09e11, 10.0, 0.25]])
        ctx.set_potential(bad)


def test_null_potential_is_free_motion(ctx):
    """gala's NullPotential (online-interface/potentials.py:103): zero components, straight lines."""
    pot = CompositePotential(names=[], types=[], params=[])
    ctx.set_potential(pot)
    rng = np.random.default_rng(1)
    n = 500
    w0 = np.vstack([rng.normal(0, 5, (3, n)), rng.normal(0, 0.2, (3, n))])
    for integ in (_lib.LEAPFROG, _lib.DOP853):
        r = ctx.integrate(w0, np.zeros(n), np.full(n, 100.0), np.ones(n), np.ones(n), np.zeros(n, dtype=np.int32),
                          integrator=integ)
        assert (r["status"] == 0).all() and (r["n_nodes"] == 100).all()
        assert np.allclose(r["w_final"][:3], w0[:3] + 99.0 * w0[3:], rtol=0, atol=1e-11)
        assert np.array_equal(r["w_final"][3:], w0[3:])


def test_pointwise_potential_and_kat1(ctx, oracle):
    v2 = MilkyWayPotential("v2")
    ctx.set_potential(v2)
    kats = json.load(open(os.path.join(HERE, "golden", "gala_notebook_kats.json")))
    vc = ctx.circular_velocity(np.array([[8.5], [0.0], [0.0]]))[0] * K.KMS_PER_KPC_MYR
    assert abs(vc - kats["KAT1_vcirc_kms_at_8p5kpc_v2"]) < 5e-9
    rng = np.random.default_rng(2)
    q = rng.normal(0, 7, (3, 4096))
    for ver in ("v1", "v2"):
        pot = MilkyWayPotential(ver)
        ctx.set_potential(pot)
        g = ctx.gradient(q)
        go = np.array([oracle.gradient(pot, q[:, i]) for i in range(q.shape[1])]).T
        assert (np.abs(g - go) / np.linalg.norm(go, axis=0)).max() < 1e-13
        v = ctx.potential_value(q)
        vo = np.array([oracle.value(pot, q[:, i]) for i in range(q.shape[1])])
        assert (np.abs(v / vo - 1)).max() < 1e-14


def test_kat5_notebook_orbit_on_gpu(ctx):
    """The six printed samples of the reference's real gala DOPRI853 orbit, reproduced by the kernel."""
    kats = json.load(open(os.path.join(HERE, "golden", "gala_notebook_kats.json")))
    ctx.set_potential(MilkyWayPotential("v2"))
    for pos, vel, t in ((kats["KAT5_pos_first3_kpc"], kats["KAT5_vel_first3_kpc_per_myr"], kats["KAT4_t_first3_myr"]),
                        (kats["KAT5_pos_last3_kpc"], kats["KAT5_vel_last3_kpc_per_myr"], kats["KAT4_t_last3_myr"])):
        w0 = np.array(pos[0] + vel[0]).reshape(6, 1)
        # seconds grid with appended t2, as the kicked notebook orbit had
        t1_s, t2_s = t[0] * K.SEC_PER_MYR, t[2] * K.SEC_PER_MYR
        r = ctx.integrate(w0, t1_s, t2_s, 0.1 * K.SEC_PER_MYR, K.MYR_PER_SEC, 1, integrator=_lib.DOP853, store_all=True)
        assert r["n_nodes"][0] == 3
        got = np.vstack([r["traj_pos"], r["traj_vel"]]).T[1:]
        exp = np.array([pos[1] + vel[1], pos[2] + vel[2]])
        assert np.abs(got - exp).max() < 2.5e-8
        assert np.allclose(r["traj_t"], t, atol=2e-8, rtol=0)


def test_math_selftest_and_peak(ctx):
    st = ctx.selftest_math(1 << 20)
    assert st["rsqrt"] < 4e-16 and st["rcp"] < 4e-16 and st["log"] < 1e-15, st
    peak = ctx.measure_fp64_peak(0, 20.0)
    assert 10.0 < peak < 60.0, peak            # B200 FP64: ~37 TFLOP/s


def test_dop853_dense_mode(ctx, oracle):
    """CW_DOP853_DENSE (one dop853 call per segment, contd8 at the grid nodes) against the
    oracle's dense restatement: bookkeeping bit-exact, same controller (step counts), states within the
    DOPRI853 tolerance; the final-state run equals the last sample of the store_all run bit for bit; and
    the dense and restart modes agree with each other to the north_star tolerance."""
    pop = make_population(300, cfg_id=8, horizon_gyr=0.4)
    g, o = run_both(ctx, oracle, pop, _lib.DOP853_DENSE)
    check_bit_exact_bookkeeping(g, o)
    assert (g["status"] == 0).mean() > 0.98
    ok = (g["status"] == 0) & (o["status"] == 0)
    d = rel_diff(g["w_final"][:, ok], o["w_final"][:, ok])
    assert np.median(d) < 1e-10 and np.quantile(d, 0.9) < 1e-7, np.quantile(d, [0.5, 0.9, 0.99, 1])
    assert abs(g["stats"][0] - o["stats"][0]) <= 5e-3 * o["stats"][0]
    if (g["status"] == 0).all():
        assert g["stats"][3] == int((g["n_nodes"] - 1).sum())
    # trajectories
    gs, os_ = run_both(ctx, oracle, pop, _lib.DOP853_DENSE, store_all=True)
    check_bit_exact_bookkeeping(gs, os_)
    off = gs["traj_off"]
    assert np.array_equal(gs["traj_t"], os_["traj_t"]) and not np.isnan(gs["traj_pos"][:, : off[-1]]).any()
    assert np.array_equal(gs["w_final"], g["w_final"])
    assert np.array_equal(gs["traj_pos"][:, off[1:] - 1][:, ok], gs["w_final"][:3][:, ok])
    assert np.array_equal(gs["traj_pos"][:, off[:-1]], pop.batch.w0[:3])
    dp = np.abs(gs["traj_pos"] - os_["traj_pos"]).max(axis=0)
    assert np.median(dp) < 1e-9 and np.quantile(dp, 0.95) < 1e-7, np.quantile(dp, [0.5, 0.95, 1])
    # the two DOPRI853 modes against each other
    r, _ = run_both(ctx, oracle, pop, _lib.DOP853)
    both = ok & (r["status"] == 0)
    dm = rel_diff(g["w_final"][:, both], r["w_final"][:, both])
    # (the dense form takes steps of several Myr: its global error after 0.4 Gyr at tol 1e-10 is ~1e-8..1e-6,
    # the restart form's forced 1 Myr steps are far more accurate than the tolerance asks)
    assert np.median(dm) < 1e-7 and np.quantile(dm, 0.9) < 2e-6, np.quantile(dm, [0.5, 0.9, 1])
    # dense needs far fewer gradient evaluations for the same tolerance
    assert g["stats"][2] < 0.6 * r["stats"][2]


def test_dop853_dense_large_batch_sorted_queue(ctx, oracle):
    """Dense mode through the sorted queue (pilot + LPT order) equals the small-batch results bit for bit."""
    pop = make_population(3000, cfg_id=9, horizon_gyr=0.2)
    b = pop.batch
    ctx.set_potential(pop.potential)
    full = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                         integrator=_lib.DOP853_DENSE, atol=1e-8, rtol=1e-8)
    assert b.n_orbits >= 4096
    idx = np.arange(0, b.n_orbits, 37)
    sub = b.take(idx)
    part = ctx.integrate(sub.w0, sub.t1, sub.t2, sub.dt, sub.to_myr, sub.flags, sub.ev_off, sub.ev_time, sub.ev_dv,
                         integrator=_lib.DOP853_DENSE, atol=1e-8, rtol=1e-8)
    assert np.array_equal(full["w_final"][:, idx], part["w_final"])
    assert np.array_equal(full["status"][idx], part["status"])


def test_store_all_ragged_lengths_and_edge_orbits(ctx, oracle):
    """store_all with every orbit a different length (Wagg2022 horizons: lanes finish at different steps, every
    column phase occurs, the warp clock delays each start by up to 15 steps) plus degenerate orbits: one- and
    two-node grids, a kick before the first node, an orbit the pre-pass rejects.  Trajectory layout, time stamps
    and kick nodes bit-exact against the oracle; samples within the leapfrog tolerance."""
    pop = make_population(160, cfg_id=21)                      # horizons 0.1 .. 12 Gyr
    b = pop.batch
    n0 = b.n_orbits
    # append hand-made edge orbits (events branch: grid in seconds, t2 appended)
    S, M = K.SEC_PER_GYR, K.SEC_PER_MYR
    w = np.array([8.0, 0.1, 0.05, 0.0, 0.22, 0.01])
    extra = [
        dict(t1=11.9999 * S, t2=12.0 * S, dt=0.1 * M, ev=[]),                         # 2 nodes (dt == span)
        dict(t1=11.9 * S, t2=12.0 * S, dt=1.0 * M, ev=[(11.8 * S, (0.01, 0.0, 0.0))]),   # kick before t1 -> node 0
        dict(t1=11.99 * S, t2=12.0 * S, dt=1.0 * M, ev=[(12.5 * S, (0.01, 0.0, 0.0))]),  # kick after t2 -> rejected
        dict(t1=12.0 * S, t2=12.0 * S, dt=1.0 * M, ev=[]),                            # empty span -> rejected
        dict(t1=11.95 * S, t2=12.0 * S, dt=1.0 * M, ev=[(11.96 * S, (0.0, 0.02, 0.0)), (11.96 * S, (0.0, 0.0, 0.02))]),
    ]
    ne = len(extra)
    w0 = np.concatenate([b.w0, np.tile(w[:, None], (1, ne))], axis=1)
    t1 = np.concatenate([b.t1, [e["t1"] for e in extra]])
    t2 = np.concatenate([b.t2, [e["t2"] for e in extra]])
    dt = np.concatenate([b.dt, [e["dt"] for e in extra]])
    to_myr = np.concatenate([b.to_myr, np.full(ne, K.MYR_PER_SEC)])
    flags = np.concatenate([b.flags, np.ones(ne, dtype=np.int32)])
    ev_off = list(b.ev_off)
    ev_time, ev_dv = list(b.ev_time), [tuple(x) for x in b.ev_dv.reshape(-1, 3)]
    for e in extra:
        for te, dv in e["ev"]:
            ev_time.append(te); ev_dv.append(dv)
        ev_off.append(len(ev_time))
    ev_off, ev_time, ev_dv = np.array(ev_off, dtype=np.int64), np.array(ev_time), np.array(ev_dv).reshape(-1, 3)
    ctx.set_potential(pop.potential)
    for integ, tol in ((_lib.LEAPFROG, 1e-9), (_lib.DOP853_DENSE, 1e-6)):
        g = ctx.integrate(w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv, integrator=integ, store_all=True)
        o = oracle.integrate_batch(pop.potential, w0, t1, t2, dt, to_myr, flags, ev_off, ev_time, ev_dv,
                                   integrator=integ, traj_off=g["traj_off"])
        assert np.array_equal(g["n_nodes"], o["n_nodes"]) and np.array_equal(g["status"], o["status"])
        assert list(g["status"][n0:]) == [0, 0, -1, -1, 0] and list(g["n_nodes"][n0:n0 + 2]) == [2, 101]
        ok = g["status"] == 0
        assert np.array_equal(g["ev_step"][: b.n_events], o["ev_step"][: b.n_events])
        assert g["ev_step"][b.n_events] == 0                     # the kick before t1 lands on node 0
        off = g["traj_off"]
        cols = np.concatenate([np.arange(off[i], off[i + 1]) for i in np.flatnonzero(ok)])
        assert np.array_equal(g["traj_t"][cols], o["traj_t"][cols])
        assert not np.isnan(g["traj_pos"][:, cols]).any()
        dpos = np.abs(g["traj_pos"][:, cols] - o["traj_pos"][:, cols]).max(axis=0)
        assert np.median(dpos) < 1e-11 if integ == _lib.LEAPFROG else np.median(dpos) < 1e-8
        assert np.quantile(dpos, 0.9) < tol * 30, np.quantile(dpos, [0.5, 0.9, 0.99, 1])
        # first sample = (kicked) initial state, last sample = w_final, for every accepted orbit
        first, last = off[:-1][ok], (off[1:] - 1)[ok]
        assert np.array_equal(g["traj_pos"][:, first], w0[:3][:, ok])
        assert np.array_equal(g["traj_pos"][:, last], g["w_final"][:3][:, ok])
        assert np.array_equal(g["traj_vel"][:, last], g["w_final"][3:][:, ok])
        # the node-0 kick is in the first stored velocity
        i = n0 + 1
        assert abs(g["traj_vel"][0, off[i]] - (w[3] + 0.01)) < 1e-15
    # an empty batch is a no-op
    e = ctx.integrate(np.zeros((6, 0)), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32),
                      integrator=_lib.LEAPFROG)
    assert e["w_final"].shape == (6, 0)


def _local_step_check(ctx, oracle, pop, label, integ=None, tol_p=1e-9, tol_v=1e-9, tol_med=1e-13):
    """EVERY stored leapfrog sample against the CPU oracle, locally: sample j+1 must be one oracle KDK step
    from sample j (no chaotic amplification, so the bound is tight and a single misplaced or stale column
    shows).  Transitions into a kick node (the stored velocity is post-kick) and into an appended t2 node
    (gala's leapfrog keeps the segment's dt there, quirk Q2) are checked by position / skipped."""
    b = pop.batch
    ctx.set_potential(pop.potential)
    integ = _lib.LEAPFROG if integ is None else integ
    g = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                      integrator=integ, store_all=True)
    assert (g["status"] == 0).all()
    off, t = g["traj_off"], g["traj_t"]
    n = b.n_orbits
    # the final-state kernel (another template instance, another queue consumption pattern) gives the same bits
    f = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ)
    assert np.array_equal(f["w_final"], g["w_final"]) and np.array_equal(f["ev_step"], g["ev_step"])
    assert np.array_equal(g["traj_pos"][:, off[1:] - 1], g["w_final"][:3])
    assert np.array_equal(g["traj_vel"][:, off[1:] - 1], g["w_final"][3:])
    lens = off[1:] - off[:-1]
    col = np.arange(off[-1])
    orbit_of = np.repeat(np.arange(n), lens)
    j = col - off[orbit_of]
    is_last = j == lens[orbit_of] - 1
    # kick nodes as absolute columns
    ev_orbit = np.repeat(np.arange(n), b.ev_off[1:] - b.ev_off[:-1])
    kick_cols = off[ev_orbit] + g["ev_step"]
    to_kick = np.zeros(off[-1] + 1, dtype=bool)
    to_kick[kick_cols] = True
    appended = (b.flags[orbit_of] & 1) == 1
    src = col[~is_last]
    dst = src + 1
    skip = appended[src] & is_last[dst] & (integ == _lib.LEAPFROG)   # the short appended interval (leapfrog only)
    src, dst = src[~skip], dst[~skip]
    w_src = np.vstack([g["traj_pos"][:, src], g["traj_vel"][:, src]])
    dt = t[dst] - t[src]
    # the segment's own dt: the first interval after the last kick node (or the orbit start) before src
    # DOPRI853 (either form): one accurate restart-form interval from sample j; the dense form's interpolated
    # samples carry its tolerance-level interpolation error, hence the looser bounds passed in
    o = oracle.integrate_batch(pop.potential, w_src, t[src], t[src] + 1.5 * dt, dt, 1.0, 0,
                               integrator=oracle.LEAPFROG if integ == _lib.LEAPFROG else oracle.DOP853)
    assert (o["n_nodes"] == 2).all() and (o["status"] == 0).all()
    dp = np.abs(o["w_final"][:3] - g["traj_pos"][:, dst]).max(axis=0)
    dv = np.abs(o["w_final"][3:] - g["traj_vel"][:, dst]).max(axis=0)
    r = np.maximum(np.linalg.norm(g["traj_pos"][:, dst], axis=0), 1e-3)
    smooth = ~to_kick[dst]
    # positions everywhere; velocities except into kick nodes.  A step through the nucleus (r << 1 kpc) amplifies the
    # 1e-13 difference between the interval and the segment's dt, hence the relative bound with a floor on r
    assert (dp / r).max() < tol_p, (label, float((dp / r).max()), int(np.argmax(dp / r)))
    assert (dv[smooth]).max() < tol_v, (label, float(dv[smooth].max()))
    assert np.median(dp) < tol_med
    return off[-1]


def test_store_all_every_sample_locally_exact(ctx, oracle):
    cols = 0
    cols += _local_step_check(ctx, oracle, make_population(200, cfg_id=21), "ragged horizons")
    cols += _local_step_check(ctx, oracle, make_population(5000, cfg_id=3, horizon_gyr=0.25), "mixed 250/251 nodes, sorted queue")
    cols += _local_step_check(ctx, oracle, make_population(400, cfg_id=4, horizon_gyr=1.0), "kicked 1001 nodes")
    assert cols > 3_000_000


def test_store_all_dop853_every_sample_locally_exact(ctx, oracle):
    """The same per-sample check for the DOPRI853 trajectories: restart form (the identical algorithm, so tight)
    and dense form (interior samples are contd8 values: bounded by the interpolation error at tol 1e-10)."""
    pop = make_population(300, cfg_id=23, horizon_gyr=0.3)
    _local_step_check(ctx, oracle, pop, "dop853 restart", integ=_lib.DOP853, tol_p=1e-10, tol_v=1e-10, tol_med=1e-13)
    _local_step_check(ctx, oracle, pop, "dop853 dense", integ=_lib.DOP853_DENSE, tol_p=3e-8, tol_v=3e-8, tol_med=1e-9)
    ragged = make_population(60, cfg_id=24)
    _local_step_check(ctx, oracle, ragged, "dop853 dense ragged", integ=_lib.DOP853_DENSE, tol_p=3e-7, tol_v=3e-7, tol_med=1e-9)
