"""CW_MEM_HOST data path (cw_api.cu integrate_host): compact cw_batch forms, page-locked vs staged buffers, chunking.

Every variant must return the SAME BITS as the plain layout in one piece: the forms only change how the bytes
reach the device (reference: the arguments `_iter_orbit_args` yields, pop.py:1156-1193, however they are shipped).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pop(n=6000, horizon=0.15, cfg=41, **kw):
    from cogsworth_b200.synthetic import make_population
    return make_population(n, cfg_id=cfg, horizon_gyr=horizon, **kw)


def _plain(ctx, b, **kw):
    return ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, **kw)


def _same(a, b, keys=("w_final", "status", "n_nodes", "ev_step")):
    for k in keys:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("integrator", [0, 1])
def test_compact_batch_equals_plain_layout(ctx, integrator):
    pop = _pop(3000 if integrator else 6000)
    b = pop.batch
    assert b.compact is not None and len(b.compact["w0_index"]) > 0
    ctx.set_potential(pop.potential)
    ref = _plain(ctx, b, integrator=integrator)
    out = ctx.integrate_batch(b, integrator=integrator)
    _same(ref, out)
    assert (ref["status"] == 0).all()


def test_compact_batch_in_many_chunks_and_forced_staging(ctx, monkeypatch):
    pop = _pop(9000, horizon=None)                    # Wagg2022 horizons: un-kicked and kicked grids mixed
    b = pop.batch
    ctx.set_potential(pop.potential)
    ref = _plain(ctx, b)
    monkeypatch.setenv("CW_HOST_CHUNK", "1500")      # ~11 chunks over the three stream lanes
    _same(ref, ctx.integrate_batch(b))
    _same(ref, _plain(ctx, b))
    monkeypatch.setenv("CW_FORCE_STAGING", "1")      # every array through the page-locked ring + host thread pool
    _same(ref, ctx.integrate_batch(b))
    _same(ref, _plain(ctx, b))


def test_pageable_and_page_locked_arrays_agree(ctx):
    from cogsworth_b200 import _lib
    pop = _pop(5000)
    b = pop.batch
    ctx.set_potential(pop.potential)
    assert _lib.is_pinned(b.compact["w0"]) and _lib.is_pinned(b.ev_time)
    assert not _lib.is_pinned(np.empty(1000))
    ref = ctx.integrate_batch(b)
    assert _lib.is_pinned(ref["w_final"])            # results land in page-locked memory directly
    c = dict(b.compact)
    for k in ("w0", "t1", "w0_index", "grid_class", "ev_off32"):
        c[k] = np.array(c[k], copy=True)             # plain numpy: pageable
    from cogsworth_b200.batch import OrbitBatch
    b2 = OrbitBatch(ev_time=np.array(b.ev_time), ev_dv=np.array(b.ev_dv), compact=c)
    _same(ref, ctx.integrate_batch(b2))


def test_store_all_compact_equals_plain(ctx, monkeypatch):
    pop = _pop(1500, horizon=0.08)
    b = pop.batch
    ctx.set_potential(pop.potential)
    ref = _plain(ctx, b, store_all=True)
    keys = ("w_final", "status", "n_nodes", "ev_step", "traj_off", "traj_pos", "traj_vel", "traj_t")
    _same(ref, ctx.integrate_batch(b, store_all=True), keys)
    monkeypatch.setenv("CW_TRAJ_SHARD_COLS", "20000")
    _same(ref, ctx.integrate_batch(b, store_all=True), keys)
    monkeypatch.setenv("CW_FORCE_STAGING", "1")
    _same(ref, ctx.integrate_batch(b, store_all=True), keys)
    _same(ref, _plain(ctx, b, store_all=True), keys)
    # and the stored trajectories end on the final states
    last = ref["traj_off"][1:] - 1
    assert np.array_equal(ref["traj_pos"][:, last], ref["w_final"][:3])


def test_count_nodes_without_phase_space(ctx):
    pop = _pop(4000, horizon=None)
    b = pop.batch
    nn, es = ctx.count_nodes(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
    nn2, es2 = ctx.count_nodes(None, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
    assert np.array_equal(nn, nn2) and np.array_equal(es, es2)
    nn3 = np.zeros(b.n_orbits, dtype=np.int32)       # one byte of grid class per orbit, no events
    ctx.count_nodes_raw(b.n_orbits, None, b.t1, None, None, None, None, None, None, None, nn3,
                        grid_class=b.compact["grid_class"], classes=b.compact["classes"])
    assert np.array_equal(nn, nn3)


def test_unsorted_w0_index_is_rejected_per_orbit(ctx):
    from cogsworth_b200.batch import OrbitBatch
    pop = _pop(3000)
    b = pop.batch
    ctx.set_potential(pop.potential)
    c = dict(b.compact)
    idx = np.array(c["w0_index"], copy=True)
    idx[5], idx[600] = idx[600], idx[5]              # out of order
    c["w0_index"] = idx
    b2 = OrbitBatch(ev_time=b.ev_time, ev_dv=b.ev_dv, compact=c)
    import os
    os.environ["CW_HOST_CHUNK"] = "500"
    try:
        out = ctx.integrate_batch(b2)
    finally:
        del os.environ["CW_HOST_CHUNK"]
    n = c["n_w0"]
    assert out["status"][n + 5] == -1 or out["status"][n + 600] == -1
    ok = np.ones(b.n_orbits, dtype=bool)
    ok[[n + 5, n + 600]] = False
    ref = ctx.integrate_batch(b)
    assert np.array_equal(out["w_final"][:, ok], ref["w_final"][:, ok])


def test_no_sync_calls_on_two_streams_share_the_scratch_safely(ctx):
    """Two un-synchronised device-mode calls on different streams: the second is ordered behind the first
    on the device, so both return what a synchronous call returns."""
    import torch
    from cogsworth_b200 import _lib
    pop = _pop(20000, horizon=0.3)
    b = pop.batch
    ctx.set_potential(pop.potential)
    dev = torch.device("cuda", 0)
    f64, i32, i64 = torch.float64, torch.int32, torch.int64
    D = dict(w0=torch.from_numpy(b.w0).to(dev), t1=torch.from_numpy(b.t1).to(dev), t2=torch.from_numpy(b.t2).to(dev),
             dt=torch.from_numpy(b.dt).to(dev), tm=torch.from_numpy(b.to_myr).to(dev),
             fl=torch.from_numpy(b.flags).to(dev), eo=torch.from_numpy(b.ev_off).to(dev),
             et=torch.from_numpy(b.ev_time).to(dev), ed=torch.from_numpy(np.ascontiguousarray(b.ev_dv)).to(dev))
    n = b.n_orbits

    def call(stream, no_sync):
        wf, st = torch.empty((6, n), dtype=f64, device=dev), torch.empty(n, dtype=i32, device=dev)
        ctx.integrate_raw(n, D["w0"], D["t1"], D["t2"], D["dt"], D["tm"], D["fl"], D["eo"], D["et"], D["ed"], wf, st,
                          mem_space=_lib.MEM_DEVICE, stream=stream.cuda_stream, no_sync=no_sync)
        return wf, st

    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ref, _ = call(s1, False)
    a, _ = call(s1, True)
    c, _ = call(s2, True)
    torch.cuda.synchronize()
    assert torch.equal(a, ref) and torch.equal(c, ref)


def test_pinned_policy_cached_never_pins_new_memory(ctx):
    """`pinned_policy("cached")` (what perform_galactic_evolution runs under): page-locked memory only when the
    library's cache already holds a block that fits, ordinary memory otherwise -- and a batch built that way integrates
    to the same bits through the staging ring."""
    from cogsworth_b200 import _lib
    n = 3_000_017
    _lib.pinned_trim()                              # whatever earlier tests left in the cache goes back to the driver
    with _lib.pinned_policy("cached"):
        a = _lib.pinned_empty(n)
        assert not _lib.is_pinned(a)                # nothing cached: ordinary memory, no cudaHostAlloc
    b = _lib.pinned_empty(n)                        # default policy: pins
    assert _lib.is_pinned(b)
    del b                                           # back to the library's cache
    with _lib.pinned_policy("cached"):
        c = _lib.pinned_empty(n)
        assert _lib.is_pinned(c)                    # the cached block is handed out
        with pytest.raises(ValueError):
            _lib.pinned_policy("sometimes")
    del a, c
    pop = _pop(4000, horizon=0.3, cfg=4)
    ctx.set_potential(pop.potential)
    ref = ctx.integrate_batch(pop.batch)
    with _lib.pinned_policy("cached"):
        from cogsworth_b200.batch import build_orbit_batch
        b2 = build_orbit_batch(pop.pos_kpc, pop.vel_kms, pop.tau_gyr, 12.0, 1.0, pop.disrupted, pop.primary_events,
                               pop.secondary_events)
        r2 = ctx.integrate_batch(b2)
    assert np.array_equal(ref["w_final"], r2["w_final"]) and np.array_equal(ref["status"], r2["status"])
    assert np.array_equal(ref["ev_step"], r2["ev_step"])
