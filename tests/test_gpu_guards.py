"""Guard bands around every output buffer: nothing is written outside them, everything promised is written inside.

compute-sanitizer is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.md), so the memcheck question is put
to the buffers themselves: each output of a device-mode call lives in the middle of a larger allocation filled with a
sentinel bit pattern (a NaN payload no arithmetic produces) and is pre-filled with the same pattern.  After the call the
bands must be intact and the buffers free of sentinels.  Host-mode calls get the same treatment on the caller's numpy
arrays (the D2H copies, the staging ring and its scatter).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SENT64 = np.array([0x7FF8DEADBEEF1234], dtype=np.uint64).view(np.float64)[0]
SENT32 = np.int32(-1163005939)          # 0xBAADF00D
BAND = 4096


def _banded_torch(torch, dev, n, dtype):
    total = n + 2 * BAND
    if dtype == torch.float64:
        full = torch.from_numpy(np.full(total, SENT64)).to(dev)
    else:
        full = torch.full((total,), int(SENT32), dtype=dtype, device=dev)
    return full, full[BAND:BAND + n]


def _check(torch, name, full, n, expect_written=True, written_mask=None):
    h = full.cpu().numpy()
    if h.dtype == np.float64:
        u = h.view(np.uint64)
        sent = np.uint64(0x7FF8DEADBEEF1234)
    else:
        u, sent = h, SENT32
    assert (u[:BAND] == sent).all() and (u[BAND + n:] == sent).all(), f"{name}: write outside the buffer"
    inner = u[BAND:BAND + n] == sent
    if written_mask is not None:
        assert not inner[written_mask].any(), f"{name}: {int(inner[written_mask].sum())} promised element(s) not written"
        assert inner[~written_mask].all(), f"{name}: element(s) written that should not have been"
    elif expect_written:
        assert not inner.any(), f"{name}: {int(inner.sum())} element(s) not written"


@pytest.mark.parametrize("integ", [0, 1, 2])
@pytest.mark.parametrize("ragged", [False, True])
def test_device_mode_outputs_between_guard_bands(ctx, integ, ragged):
    import torch
    from cogsworth_b200 import _lib
    from cogsworth_b200.synthetic import make_population
    pop = make_population(1300 if integ else 2600, cfg_id=61 + integ, horizon_gyr=(0.004, 0.09) if ragged else 0.07)
    b = pop.batch
    n, K = b.n_orbits, b.n_events
    ctx.set_potential(pop.potential)
    dev = torch.device("cuda", 0)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    D = [to(x) for x in (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)]
    nn, _ = ctx.count_nodes(None, b.t1, b.t2, b.dt, b.to_myr, b.flags)
    slack = np.where(np.arange(n) % 7 == 0, 3, 0)              # some trajectory slots larger than the orbit needs
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nn + slack, out=off[1:])
    tot = int(off[-1])
    f64, i32 = torch.float64, torch.int32
    wf_f, wf = _banded_torch(torch, dev, 6 * n, f64)
    st_f, st = _banded_torch(torch, dev, n, i32)
    nn_f, nnd = _banded_torch(torch, dev, n, i32)
    es_f, es = _banded_torch(torch, dev, K, i32)
    tp_f, tp = _banded_torch(torch, dev, 3 * tot, f64)
    tv_f, tv = _banded_torch(torch, dev, 3 * tot, f64)
    tt_f, tt = _banded_torch(torch, dev, tot, f64)
    ctx.integrate_raw(n, *D, wf, st, nnd, es, to(off), tp, tv, tt, integrator=integ, mem_space=_lib.MEM_DEVICE)
    torch.cuda.synchronize()
    assert (st == 0).all()
    # columns of the trajectory rows that belong to a sample (the slack columns must stay untouched)
    used = np.zeros(tot, dtype=bool)
    from cogsworth_b200.kicks import _ranges
    used[_ranges(off[:-1], nn)] = True
    _check(torch, "w_final", wf_f, 6 * n)
    _check(torch, "status", st_f, n)
    _check(torch, "n_nodes", nn_f, n)
    _check(torch, "ev_step", es_f, K)
    _check(torch, "traj_pos", tp_f, 3 * tot, written_mask=np.tile(used, 3))
    _check(torch, "traj_vel", tv_f, 3 * tot, written_mask=np.tile(used, 3))
    _check(torch, "traj_t", tt_f, tot, written_mask=used)
    assert np.array_equal(nnd.cpu().numpy(), nn)


@pytest.mark.parametrize("staged", [False, True])
def test_host_mode_outputs_between_guard_bands(ctx, staged, monkeypatch):
    from cogsworth_b200 import _lib
    from cogsworth_b200.synthetic import make_population
    pop = make_population(3000, cfg_id=66, horizon_gyr=(0.004, 0.09))
    b = pop.batch
    n, K = b.n_orbits, b.n_events
    ctx.set_potential(pop.potential)
    monkeypatch.setenv("CW_HOST_CHUNK", "700")
    monkeypatch.setenv("CW_TRAJ_SHARD_COLS", "30000")
    if staged:
        monkeypatch.setenv("CW_FORCE_STAGING", "1")
    nn, _ = ctx.count_nodes(None, b.t1, b.t2, b.dt, b.to_myr, b.flags)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nn, out=off[1:])
    tot = int(off[-1])

    def banded(m, dtype):
        alloc = _lib.pinned_empty if not staged else np.empty
        full = alloc(m + 2 * BAND, dtype)
        full[...] = SENT64 if dtype == np.float64 else SENT32
        return full, full[BAND:BAND + m]

    def check(name, full, m):
        u = full.view(np.uint64) if full.dtype == np.float64 else full
        sent = np.uint64(0x7FF8DEADBEEF1234) if full.dtype == np.float64 else SENT32
        assert (u[:BAND] == sent).all() and (u[BAND + m:] == sent).all(), f"{name}: write outside the buffer"
        assert not (u[BAND:BAND + m] == sent).any(), f"{name}: element(s) not written"

    wf_f, wf = banded(6 * n, np.float64)
    st_f, st = banded(n, np.int32)
    nn_f, nno = banded(n, np.int32)
    es_f, es = banded(K, np.int32)
    tp_f, tp = banded(3 * tot, np.float64)
    tv_f, tv = banded(3 * tot, np.float64)
    tt_f, tt = banded(tot, np.float64)
    a = b.abi_arrays()
    ctx.integrate_raw(n, a["w0"], a["t1"], None, None, None, None, None, a["ev_time"], a["ev_dv"], wf, st, nno, es, off, tp,
                      tv, tt, n_w0=a["n_w0"], w0_index=a["w0_index"], grid_class=a["grid_class"], classes=a["classes"],
                      ev_off32=a["ev_off32"])
    assert (st == 0).all()
    for name, full, m in (("w_final", wf_f, 6 * n), ("status", st_f, n), ("n_nodes", nn_f, n), ("ev_step", es_f, K),
                          ("traj_pos", tp_f, 3 * tot), ("traj_vel", tv_f, 3 * tot), ("traj_t", tt_f, tot)):
        check(name, full, m)
    # and the content is what the plain, one-piece call returns
    monkeypatch.delenv("CW_HOST_CHUNK")
    monkeypatch.delenv("CW_TRAJ_SHARD_COLS")
    monkeypatch.delenv("CW_FORCE_STAGING", raising=False)
    ref = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, store_all=True)
    assert np.array_equal(wf.reshape(6, n), ref["w_final"]) and np.array_equal(tp.reshape(3, tot), ref["traj_pos"])
    assert np.array_equal(tt, ref["traj_t"]) and np.array_equal(es, ref["ev_step"])
