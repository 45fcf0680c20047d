"""Edge cases of the drop-in boundary on the GPU: empty and one-orbit batches, degenerate spans, and arguments the C ABI
must refuse (the reference's counterparts: an empty population skips the pool, `pop.py:1239-1266`; a span shorter than
dt is clipped, `kicks.py:97`; un-stitchable orbits raise, `kicks.py:187-189`)."""
import ctypes as C

import numpy as np
import pytest

from cogsworth_b200 import _lib
from cogsworth_b200.potential import MilkyWayPotential

pytestmark = pytest.mark.gpu
E_INVALID, E_UNSUPPORTED = -1, -5          # CW_E_INVALID, CW_E_UNSUPPORTED (include/cogsworth_b200.h)


@pytest.mark.parametrize("integ", [_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE])
@pytest.mark.parametrize("store_all", [False, True])
def test_empty_batch_is_a_no_op(ctx, integ, store_all):
    ctx.set_potential(MilkyWayPotential("v2"))
    r = ctx.integrate(np.zeros((6, 0)), np.zeros(0), np.zeros(0), np.zeros(0), integrator=integ, store_all=store_all)
    assert r["w_final"].shape == (6, 0) and len(r["status"]) == 0 and len(r["n_nodes"]) == 0
    if store_all:
        assert r["traj_off"].tolist() == [0] and r["traj_pos"].shape[1] == 0


@pytest.mark.parametrize("integ", [_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE])
def test_single_orbit_equals_its_place_in_a_batch(ctx, integ):
    """One orbit on its own (a grid of one CTA, nothing to sort, no chunks) returns the bits it returns inside 5000."""
    ctx.set_potential(MilkyWayPotential("v2"))
    rng = np.random.default_rng(8)
    n = 5000
    w0 = np.vstack([rng.normal(0, 6, (3, n)), rng.normal(0, 0.15, (3, n))])
    t2 = rng.uniform(50.0, 300.0, n)
    big = ctx.integrate(w0, np.zeros(n), t2, np.ones(n), integrator=integ)
    for i in (0, 1234, n - 1):
        one = ctx.integrate(w0[:, i:i + 1].copy(), np.zeros(1), t2[i:i + 1].copy(), np.ones(1), integrator=integ)
        assert one["status"][0] == big["status"][i] == 0 and one["n_nodes"][0] == big["n_nodes"][i]
        assert np.array_equal(one["w_final"][:, 0], big["w_final"][:, i])


def test_degenerate_spans(ctx, oracle):
    """t2 == t1 (an EMPTY grid, as np.arange(t1, t2, dt): the orbit is reported inconsistent, status -1, nothing else is
    touched), a span shorter than dt with t2 appended (one short step), and a span of exactly one step -- node counts,
    statuses and states as the oracle's."""
    pot = MilkyWayPotential("v2")
    ctx.set_potential(pot)
    w0 = np.array([[8.0, 0.0, 0.1, 0.0, 0.22, 0.01]] * 3).T.copy()
    t1 = np.array([5.0, 5.0, 5.0])
    t2 = np.array([5.0, 5.4, 6.0])
    dt = np.array([1.0, 0.4, 1.0])                       # the caller has applied dt = min(dt, t2 - t1) (kicks.py:97)
    flags = np.array([1, 1, 1], dtype=np.int32)
    for integ, ointeg in ((_lib.LEAPFROG, oracle.LEAPFROG), (_lib.DOP853, oracle.DOP853)):
        g = ctx.integrate(w0, t1, t2, dt, flags=flags, integrator=integ, store_all=True)
        o = oracle.integrate_batch(pot, w0, t1, t2, dt, np.ones(3), flags, None, None, None, integrator=ointeg)
        assert np.array_equal(g["n_nodes"], o["n_nodes"]) and np.array_equal(g["status"], o["status"])
        assert g["status"].tolist() == [-1, 0, 0] and g["n_nodes"][0] == 0 and g["n_nodes"][1] >= 2
        ok = g["status"] == 0
        assert np.allclose(g["w_final"][:, ok], o["w_final"][:, ok], rtol=1e-12, atol=1e-15)
        off = g["traj_off"]
        assert np.array_equal(g["traj_pos"][:, off[:-1][ok]], w0[:3, ok])   # first stored sample = the initial state
        assert np.array_equal(g["traj_pos"][:, off[1:][ok] - 1], g["w_final"][:3, ok])


def test_abi_refuses_bad_arguments(ctx):
    """cw_integrate returns CW_E_INVALID / CW_E_UNSUPPORTED (and says why) instead of touching memory."""
    L = ctx._L
    ctx.set_potential(MilkyWayPotential("v2"))
    n = 4
    w0, t1 = np.zeros((6, n)), np.zeros(n)
    t2, dt, tm = np.full(n, 10.0), np.ones(n), np.ones(n)
    fl = np.zeros(n, dtype=np.int32)
    wf, st = np.zeros((6, n)), np.zeros(n, dtype=np.int32)

    def call(n_orbits=n, integrator=_lib.LEAPFROG, mem_space=0, t2_arr=t2, **batch_kw):
        o = _lib.CwOpts(integrator=integrator, mem_space=mem_space, atol=1e-10, rtol=1e-10, nmax=0)
        b = _lib.CwBatch(n_orbits=n_orbits, w0=w0.ctypes.data, t1=t1.ctypes.data,
                         t2=None if t2_arr is None else t2_arr.ctypes.data, dt=dt.ctypes.data, to_myr=tm.ctypes.data,
                         grid_flags=fl.ctypes.data, **batch_kw)
        r = _lib.CwResult(w_final=wf.ctypes.data, status=st.ctypes.data)
        return L.cw_integrate(ctx._h, C.byref(o), C.byref(b), C.byref(r))

    assert call() == 0
    assert call(n_orbits=-1) == E_INVALID and b"n_orbits" in L.cw_last_error(ctx._h)
    assert call(t2_arr=None) == E_INVALID                       # a null grid column
    assert call(integrator=7) == E_UNSUPPORTED
    assert call(mem_space=5) == E_INVALID
    assert call(n_w0=2) == E_INVALID                            # n_w0 < n_orbits without w0_index
    assert call(n_w0=n + 1) == E_INVALID
    # a dt of zero is an orbit-level error (the pre-pass rejects the orbit, the call succeeds)
    dt[1] = 0.0
    assert call() == 0 and st[1] == -1 and st[0] == 0
    dt[1] = 1.0
