"""The oracle's kicks.py driver (cwo_integrate_orbit) against a literal numpy transcription of the
reference loop (masks over the full time array, segment by segment), plus the structural properties
the reference's own tests assert (tests/test_kicks.py, tests/test_pop.py:232-247)."""
import numpy as np
import pytest

from cogsworth_b200 import constants as K
from cogsworth_b200.potential import MilkyWayPotential

V2 = MilkyWayPotential("v2")


def literal_driver(O, w0, t1_s, t2_s, dt_s, ev_time_s, ev_dv, integrator="leapfrog"):
    """kicks.py:113-189 transcribed with numpy masks; uses the oracle only for single segments."""
    T = [t1_s]
    # parse_time_specification: sequential accumulation
    T = []
    x = t1_s
    while x < t2_s and len(T) < 1e6:
        T.append(x)
        x += dt_s
    T = np.array(T)
    if T[-1] < t2_s:
        T = np.append(T, t2_s)
    cursor = T[0]
    w = np.array(w0, dtype=float)
    data, steps = [], []
    for te, dv in zip(ev_time_s, ev_dv):
        mask = (T >= cursor) & (T < te)
        if mask.any():
            seg = T[mask] * K.MYR_PER_SEC
            traj = O.leapfrog(V2, w, seg) if len(seg) > 1 else w.reshape(6, 1).copy()
            data.append(traj[:, :-1])
            w = traj[:, -1].copy()
            cursor = T[mask][-1]
            steps.append(int(np.flatnonzero(mask)[-1]))
        else:
            cursor = te
            steps.append(sum(d.shape[1] for d in data))
        w[3:] += dv
    if cursor < T[-1]:
        seg = T[T >= cursor] * K.MYR_PER_SEC
        data.append(O.leapfrog(V2, w, seg))
    return np.concatenate(data, axis=1), T * K.MYR_PER_SEC, steps


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_driver_matches_literal_transcription(oracle, seed):
    rng = np.random.default_rng(seed)
    w0 = np.array([rng.uniform(4, 10), rng.uniform(-2, 2), rng.uniform(-0.3, 0.3), 0.01, 0.21, 0.01])
    t1_gyr = 12.0 - rng.uniform(0.05, 0.3)
    t1, t2, dt = t1_gyr * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR, 1.0 * K.SEC_PER_MYR
    tphys = np.sort(rng.uniform(0.0, 40.0, 3))
    if seed == 1:
        tphys[1] = tphys[0] + 0.2          # two supernovae inside one grid interval
    if seed == 2:
        tphys[0] = 0.0                     # event at birth: empty mask branch (kicks.py:155-157)
    ev_time = (t1_gyr + tphys * K.GYR_PER_MYR) * K.SEC_PER_GYR
    ev_dv = rng.normal(0, 0.1, (3, 3))
    r = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, ev_time, ev_dv, integrator=oracle.LEAPFROG,
                               store_all=True)
    traj, t_myr, steps = literal_driver(oracle, w0, t1, t2, dt, ev_time, ev_dv)
    assert r["status"] == 0
    assert list(r["ev_step"]) == steps
    assert r["traj"].shape == traj.shape
    assert np.array_equal(r["traj"], traj)          # same arithmetic, same order: bit-identical
    assert np.array_equal(r["t"], t_myr)
    assert np.array_equal(r["w_final"], traj[:, -1])


def test_no_duplicated_stitch_samples(oracle):
    """tests/test_kicks.py:7-38: diff(t), diff(x|y|z) never 0 for kicked orbits."""
    w0 = np.array([8.0, 0.0, 0.05, 0.0, 0.22, 0.0])
    t1, t2, dt = 11.8 * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR, K.SEC_PER_MYR
    ev_time = (11.8 + np.array([10.3, 55.7]) * K.GYR_PER_MYR) * K.SEC_PER_GYR
    r = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, ev_time, [[0.05, 0, 0.02], [0, -0.1, 0]],
                               integrator=oracle.DOP853, store_all=True)
    assert r["status"] == 0 and r["traj"].shape[1] == len(r["t"]) == r["n_nodes"]
    assert (np.diff(r["t"]) != 0).all()
    for c in range(3):
        assert (np.diff(r["traj"][c]) != 0).all()
    # the sample stored ON the kick node is the post-kick state: velocity jumps exactly there
    k = r["ev_step"][0]
    no_kick = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, integrator=oracle.DOP853, store_all=True)
    assert np.array_equal(r["traj"][:, :k], no_kick["traj"][:, :k])
    assert np.allclose(r["traj"][3:, k] - no_kick["traj"][3:, k], [0.05, 0, 0.02], atol=1e-15)
    assert np.array_equal(r["traj"][:3, k], no_kick["traj"][:3, k])


def test_store_all_false_is_last_sample(oracle):
    """tests/test_pop.py:232-247: store_entire_orbits=False keeps orbit[-1:]."""
    w0 = np.array([6.0, 1.0, 0.2, -0.02, 0.2, 0.01])
    args = (V2, w0, 0.0, 200.0, 1.0, 1.0, 0)
    full = oracle.integrate_orbit(*args, integrator=oracle.LEAPFROG, store_all=True)
    last = oracle.integrate_orbit(*args, integrator=oracle.LEAPFROG, store_all=False)
    assert np.array_equal(full["traj"][:, -1], last["w_final"])
    # un-kicked branch: gala's own grid stops at the last node < t2 (quirk Q1): 200 nodes, t = 0..199
    assert full["n_nodes"] == 200 and full["t"][-1] == 199.0


def test_event_after_t2_is_inconsistent(oracle):
    """A kick on the final node leaves no tail segment; the reference raises when stitching."""
    w0 = np.array([8.0, 0.0, 0.0, 0.0, 0.22, 0.0])
    t1, t2, dt = 11.9 * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR, K.SEC_PER_MYR
    r = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, [t2 * 1.001], [[0.1, 0, 0]],
                               integrator=oracle.LEAPFROG)
    assert r["status"] == -1


def test_secondary_prefix_bit_identical(oracle):
    """plot.py:558 compares primary and secondary xyz exactly before the disruption."""
    w0 = np.array([5.0, -3.0, 0.1, 0.1, 0.15, -0.01])
    t1g = 11.7
    t1, t2, dt = t1g * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR, K.SEC_PER_MYR
    te = (t1g + np.array([20.0, 33.3]) * K.GYR_PER_MYR) * K.SEC_PER_GYR
    prim = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, te, [[0.02, 0.01, 0], [0.03, 0, 0.01]],
                                  integrator=oracle.DOP853, store_all=True)
    sec = oracle.integrate_orbit(V2, w0, t1, t2, dt, K.MYR_PER_SEC, 1, te, [[0.02, 0.01, 0], [-0.2, 0.1, 0.1]],
                                 integrator=oracle.DOP853, store_all=True)
    k = prim["ev_step"][1]
    assert np.array_equal(prim["traj"][:, :k], sec["traj"][:, :k])
    assert not np.array_equal(prim["traj"][:, k:], sec["traj"][:, k:])


def test_batch_equals_single_and_threads(oracle):
    from cogsworth_b200.synthetic import make_population
    pop = make_population(40, cfg_id=7, horizon_gyr=0.1)
    b = pop.batch
    r1 = oracle.integrate_batch(V2, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                                integrator=oracle.LEAPFROG, n_threads=1)
    r4 = oracle.integrate_batch(V2, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                                integrator=oracle.LEAPFROG, n_threads=4)
    assert np.array_equal(r1["w_final"], r4["w_final"]) and np.array_equal(r1["ev_step"], r4["ev_step"])
    for i in (0, 17, b.n_orbits - 1):
        e0, e1 = b.ev_off[i], b.ev_off[i + 1]
        s = oracle.integrate_orbit(V2, b.w0[:, i], b.t1[i], b.t2[i], b.dt[i], b.to_myr[i], int(b.flags[i]),
                                   b.ev_time[e0:e1], b.ev_dv[e0:e1], integrator=oracle.LEAPFROG)
        assert np.array_equal(s["w_final"], r1["w_final"][:, i])
        assert list(s["ev_step"]) == list(r1["ev_step"][e0:e1])


def test_backward_grid_and_rewind_roundtrip(oracle):
    """Backward integration (cogsworth hydro/rewind.py:16-19): gala's backward grid -- nodes
    while T > t2, then t2 itself -- and forward-then-backward round trips with both integrators."""
    from cogsworth_b200.potential import MilkyWayPotential
    T = oracle.time_grid(1000.0, 990.5, -1.0, True)
    assert len(T) == 11 and T[0] == 1000.0 and T[-2] == 991.0 and T[-1] == 990.5
    T = oracle.time_grid(1000.0, 990.0, -1.0, True)
    assert len(T) == 11 and T[-1] == 990.0 and T[-2] == 991.0       # t2 is appended, never duplicated
    assert len(oracle.time_grid(1000.0, 990.0, -1.0, False)) == 10
    assert len(oracle.time_grid(0.0, 10.0, -1.0, True)) == 0         # dt < 0 with t2 > t1: gala raises
    assert len(oracle.time_grid(10.0, 0.0, 1.0, True)) == 0          # dt > 0 with t2 < t1: gala raises
    v2 = MilkyWayPotential("v2")
    w = np.array([8.0, 0.3, 0.1, 0.01, 0.22, 0.02])
    for integ, tol in ((oracle.LEAPFROG, 1e-13), (oracle.DOP853, 1e-11), (oracle.DOP853_DENSE, 1e-6)):   # dense: global error of the larger steps
        f = oracle.integrate_orbit(v2, w, 0.0, 300.0, 1.0, integrator=integ)          # ends on node 299
        b = oracle.integrate_orbit(v2, f["w_final"], 299.0, 0.0, -1.0, flags=1, integrator=integ)
        assert b["status"] == 0 and b["n_nodes"] == 300
        assert np.abs(b["w_final"] - w).max() < tol, (integ, np.abs(b["w_final"] - w).max())
    # events never occur on a backward grid
    r = oracle.integrate_orbit(v2, w, 299.0, 0.0, -1.0, flags=1, ev_time=[100.0], ev_dv=[[0.01, 0.0, 0.0]])
    assert r["status"] == -1
