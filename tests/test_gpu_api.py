"""GPU tests of the reference-facing API (per-orbit call, Population stage, retry logic, device
pointers) and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200 import constants as K
from cogsworth_b200.potential import MilkyWayPotential
from cogsworth_b200.synthetic import make_population, population_with_n_orbits

pytestmark = pytest.mark.gpu


def test_per_orbit_api_matches_batch_and_oracle(ctx, oracle):
    """integrate_orbit_with_events(w0, t1, t2, dt, potential, events, ...) -- reference signature."""
    import pandas as pd
    pot = MilkyWayPotential("v2")
    w0 = [8.0, 0.3, 0.05, 5.0, 215.0, 3.0]                       # kpc, km/s
    events = pd.DataFrame({"tphys": [12.3, 40.1], "delta_vsys_x": [30., -80.], "delta_vsys_y": [10., 50.],
                           "delta_vsys_z": [-5., 120.], "phase": [0.3, 2.0], "inc": [1.1, 0.4]})
    for integ, O_int in (("leapfrog", oracle.LEAPFROG), (None, oracle.DOP853)):
        orb = cb.integrate_orbit_with_events(w0, 11.8, 12.0, 1.0, potential=pot, events=events, store_all=True,
                                             integrator=integ, context=ctx)
        assert orb.shape[0] in (201, 202) and orb.t[-1] == 12000.0
        dv = cb.batch.rotate_kicks(events[["delta_vsys_x", "delta_vsys_y", "delta_vsys_z"]].values,
                                   events["phase"].values, events["inc"].values) * K.KPC_MYR_PER_KMS
        w_gal = np.array(w0[:3] + [v * K.KPC_MYR_PER_KMS for v in w0[3:]])
        te = (11.8 + events["tphys"].values * K.GYR_PER_MYR) * K.SEC_PER_GYR
        ref = oracle.integrate_orbit(pot, w_gal, 11.8 * K.SEC_PER_GYR, 12.0 * K.SEC_PER_GYR, K.SEC_PER_MYR,
                                     K.MYR_PER_SEC, 1, te, dv, integrator=O_int, store_all=True)
        assert orb.shape[0] == ref["n_nodes"] and np.array_equal(orb.t, ref["t"])
        assert np.abs(orb.pos.xyz - ref["traj"][:3]).max() < 1e-9
        last = cb.integrate_orbit_with_events(w0, 11.8, 12.0, 1.0, potential=pot, events=events, store_all=False,
                                              integrator=integ, context=ctx)
        assert last.shape == (1,) and np.array_equal(last.pos.xyz[:, 0], orb.pos.xyz[:, -1])
    # no events: gala's own grid, which stops at the last node < t2 (quirk Q1)
    orb = cb.integrate_orbit_with_events(w0, 11.8, 12.0, 1.0, potential=pot, events=None, integrator="leapfrog",
                                         context=ctx)
    assert orb.shape == (200,) and abs(orb.t[-1] - 11999.0) < 1e-9
    # dt larger than the span is clipped: dt = min(dt, t2 - t1)  (kicks.py:97)
    orb = cb.integrate_orbit_with_events(w0, 11.9995, 12.0, 1.0, potential=pot, events=events.iloc[:0],
                                         integrator="leapfrog", context=ctx)
    assert orb.shape[0] in (2, 3) and orb.t[-1] == 12000.0


def test_population_stage_end_to_end(ctx):
    """B200Population.perform_galactic_evolution with COSMIC-style tables (reference tests/test_pop.py
    :232-247 store_entire_orbits, :841-864 primary/secondary views, tests/test_kicks.py:40-52 rerun)."""
    from test_host import fake_population
    p = fake_population()
    p._cw_context = ctx
    np.random.seed(11)
    p.perform_galactic_evolution(progress_bar=False)
    n, nd = len(p), int(p.disrupted.sum())
    assert len(p.orbits) == n + nd == 6
    assert all(o.shape[0] == 1 for o in p.orbits)               # store_entire_orbits=False
    assert p.final_pos.shape == (6, 3) and p.final_vel.shape == (6, 3)
    assert p.final_primary_pos.shape == (4, 3) and p.final_secondary_pos.shape == (4, 3)
    # bound systems: secondary == primary; disrupted: the extra orbits
    assert np.array_equal(p.final_secondary_pos[:2], p.final_primary_pos[:2])
    assert np.array_equal(p.final_secondary_pos[2:], p.final_pos[4:])
    first = np.array(p.final_pos).copy()
    p.perform_galactic_evolution(progress_bar=False)            # angles persisted -> reproducible
    assert np.array_equal(np.array(p.final_pos), first)
    # full orbits + DOPRI853 default
    p.store_entire_orbits, p.integrator = True, None
    p.perform_galactic_evolution()
    lens = [o.shape[0] for o in p.orbits]
    assert lens[0] == 1000 and lens[1] == 2001 and lens[3] == 4001 and lens[5] == 4001
    assert np.array_equal(np.array(p.primary_orbits[1].pos.xyz)[:, -1], np.array(p.final_pos)[1])


def test_retry_with_smaller_timestep(ctx, oracle):
    """kicks.py:113,193-203: failed kicked orbits are retried with dt * timestep_multiplier,
    max_retries attempts in total; un-kicked ones are not retried; exhausted -> None."""
    pop = make_population(200, cfg_id=21, horizon_gyr=0.1, plunging_fraction=0.5)
    b = pop.batch
    kw = dict(potential=pop.potential, store_all=False, integrator="dopri853", context=ctx)
    orbits0, info0 = cb.integrate_orbits(b, integrator_kwargs={"nmax": 1}, max_retries=1, **kw)
    failed0 = np.flatnonzero(info0["status"] != 0)
    assert len(failed0) > 0 and (info0["status"][failed0] == -2).all()      # "Larger nmax is needed"
    assert all(orbits0[int(i)] is None for i in failed0[:5])
    orbits1, info1 = cb.integrate_orbits(b, integrator_kwargs={"nmax": 1}, max_retries=2, timestep_multiplier=0.1, **kw)
    recovered = failed0[info1["status"][failed0] == 0]
    assert len(recovered) > 0
    i = int(recovered[0])
    e0, e1 = b.ev_off[i], b.ev_off[i + 1]
    ref = oracle.integrate_orbit(pop.potential, b.w0[:, i], b.t1[i], b.t2[i], b.dt[i] * 0.1, b.to_myr[i],
                                 int(b.flags[i]), b.ev_time[e0:e1], b.ev_dv[e0:e1], integrator=oracle.DOP853, nmax=1)
    assert ref["status"] == 0 and info1["n_nodes"][i] == ref["n_nodes"]
    assert np.abs(info1["w_final"][:, i] - ref["w_final"]).max() < 1e-7
    fp, fv = orbits1.final_coords()
    still = np.flatnonzero(info1["status"] != 0)
    assert np.isinf(fp[still]).all() and np.isfinite(fp[info1["status"] == 0]).all()


def test_device_pointer_mode_matches_host_mode(ctx):
    import torch
    pop = make_population(3000, cfg_id=31, horizon_gyr=0.15)
    b = pop.batch
    ctx.set_potential(pop.potential)
    host = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG)
    dev = torch.device("cuda:0")
    tt = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
    d = dict(w0=tt(b.w0), t1=tt(b.t1), t2=tt(b.t2), dt=tt(b.dt), to_myr=tt(b.to_myr), flags=tt(b.flags, torch.int32),
             ev_off=tt(b.ev_off, torch.int64), ev_time=tt(b.ev_time), ev_dv=tt(b.ev_dv))
    n = b.n_orbits
    wf = torch.empty((6, n), dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    es = torch.empty(b.n_events, dtype=torch.int32, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctx.integrate_raw(n, d["w0"], d["t1"], d["t2"], d["dt"], d["to_myr"], d["flags"], d["ev_off"], d["ev_time"],
                      d["ev_dv"], wf, st, None, es, stats=stats, integrator=_lib.LEAPFROG, mem_space=_lib.MEM_DEVICE)
    assert np.array_equal(wf.cpu().numpy(), host["w_final"])
    assert np.array_equal(es.cpu().numpy(), host["ev_step"]) and int(stats[3]) == int(host["stats"][3])
    assert ctx.last_kernel_ms(0) > 0


def test_properties_at_full_size(ctx):
    """10^6 orbits x 1000 steps (a tenth of configs[3]): determinism, exact step accounting, time
    reversibility of the leapfrog map, bounded energy error -- no oracle needed."""
    pop = population_with_n_orbits(1_000_000, cfg_id=4, horizon_gyr=1.0)
    b = pop.batch
    ctx.set_potential(pop.potential)
    run = lambda w0, ev=True: ctx.integrate(w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off if ev else None,
                                            b.ev_time if ev else None, b.ev_dv if ev else None, integrator=_lib.LEAPFROG)
    r1, r2 = run(b.w0), run(b.w0)
    assert np.array_equal(r1["w_final"], r2["w_final"])                      # deterministic
    assert (r1["status"] == 0).all() and (r1["n_nodes"] == 1001).all()
    assert int(r1["stats"][3]) == 1000 * b.n_orbits                          # every orbit-step accounted for
    nn, es = ctx.count_nodes(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
    assert np.array_equal(nn, r1["n_nodes"]) and np.array_equal(es, r1["ev_step"])
    assert es.min() >= 3 and es.max() <= 70                                   # SN at 3..70 Myr after birth
    assert np.isfinite(r1["w_final"]).all()
    # reversibility (no kicks): forward, flip v, forward again, flip -> back at the start
    fwd = run(b.w0, ev=False)["w_final"]
    fwd[3:] *= -1
    back = run(fwd, ev=False)["w_final"]
    back[3:] *= -1
    err = np.abs(back - b.w0).max(axis=0) / np.maximum(np.linalg.norm(b.w0[:3], axis=0), 1e-2)
    assert np.median(err) < 1e-12 and np.quantile(err, 0.9) < 1e-9, np.quantile(err, [0.5, 0.9, 0.99])
    # energy: symplectic, so bounded for resolved orbits
    E0 = 0.5 * (b.w0[3:] ** 2).sum(axis=0) + ctx.potential_value(np.ascontiguousarray(b.w0[:3]))
    E1 = 0.5 * (fwd[3:] ** 2).sum(axis=0) + ctx.potential_value(np.ascontiguousarray(fwd[:3]))
    assert np.median(np.abs(E1 / E0 - 1)) < 1e-5


def test_in_process_multi_device_sharding(built):
    """cw_create over every visible GPU: orbits are sharded by index across the devices inside one call
    (no collective); the result equals the single-device run bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    pop = make_population(5000, cfg_id=51, horizon_gyr=0.1)
    b = pop.batch
    one = cb.Context(devices=[0])
    many = cb.Context()                       # all devices
    assert many.n_devices == torch.cuda.device_count()
    outs = []
    for c in (one, many):
        c.set_potential(pop.potential)
        outs.append(c.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                                integrator=_lib.LEAPFROG, store_all=True))
    assert np.array_equal(outs[0]["w_final"], outs[1]["w_final"])
    assert np.array_equal(outs[0]["ev_step"], outs[1]["ev_step"])
    assert np.array_equal(outs[0]["traj_pos"], outs[1]["traj_pos"])
    assert np.array_equal(outs[0]["traj_t"], outs[1]["traj_t"])
    assert int(outs[1]["stats"][3]) == int(outs[0]["stats"][3])
    # The cut over the devices balances WORK, not orbit counts: a population sorted by age (grid steps grow along the
    # index, horizons 0.02 .. 0.6 Gyr) through the compact form -- same bits as one device, and no device is left
    # with much more than its share of the kernel time.
    from cogsworth_b200.batch import build_orbit_batch
    srt = make_population(60_000, cfg_id=52, horizon_gyr=(0.02, 0.6))
    o = np.argsort(srt.tau_gyr, kind="stable")
    sb = build_orbit_batch(srt.pos_kpc[:, o], srt.vel_kms[:, o], srt.tau_gyr[o], 12.0, 1.0, np.zeros(len(o), dtype=bool))
    res = []
    for c in (one, many):
        c.set_potential(srt.potential)
        res.append(c.integrate_batch(sb))
        res.append(c.integrate_batch(sb))          # (second call: buffers are in place, kernel times are clean)
    assert np.array_equal(res[1]["w_final"], res[3]["w_final"]) and np.array_equal(res[1]["n_nodes"], res[3]["n_nodes"])
    ms = [many.last_kernel_ms(k) for k in range(many.n_devices)]
    assert max(ms) < 1.35 * (sum(ms) / len(ms)), ms      # an equal-count cut gives the last of two devices 1.5x the mean
    one.close(); many.close()


def test_several_slots_on_one_device(built):
    """The in-process multi-device path -- work-balanced contiguous cut, one long-lived host thread per slot, per-slot
    chunk pipelines -- exercised on ONE GPU by naming it several times (cw_create accepts a device more than once: the
    slots are independent): every slot count returns the single-slot bits, for page-locked compact batches in several
    chunks per slot and for an age-sorted population whose slices then differ in length."""
    from cogsworth_b200.batch import build_orbit_batch
    pop = make_population(150_000, cfg_id=53, horizon_gyr=0.05)
    srt = make_population(40_000, cfg_id=54, horizon_gyr=(0.01, 0.2))
    o = np.argsort(srt.tau_gyr, kind="stable")
    sb = build_orbit_batch(srt.pos_kpc[:, o], srt.vel_kms[:, o], srt.tau_gyr[o], 12.0, 1.0, np.zeros(len(o), dtype=bool))
    one = cb.Context(devices=[0])
    one.set_potential(pop.potential)
    ref, ref_s = one.integrate_batch(pop.batch), one.integrate_batch(sb)
    for ns in (2, 3, 8):
        many = cb.Context(devices=[0] * ns)
        assert many.n_devices == ns
        many.set_potential(pop.potential)
        for _ in range(2):                               # (the second call re-uses the slot threads and buffers)
            r = many.integrate_batch(pop.batch)
            for k in ("w_final", "status", "n_nodes", "ev_step"):
                assert np.array_equal(r[k], ref[k]), (ns, k)
            assert int(r["stats"][3]) == int(ref["stats"][3])
        rs = many.integrate_batch(sb)
        assert np.array_equal(rs["w_final"], ref_s["w_final"]) and np.array_equal(rs["n_nodes"], ref_s["n_nodes"])
        tr = many.integrate_batch(sb, store_all=True)
        tr1 = one.integrate_batch(sb, store_all=True)
        assert np.array_equal(tr["traj_pos"], tr1["traj_pos"]) and np.array_equal(tr["traj_off"], tr1["traj_off"])
        many.close()
    one.close()


def test_rewind_backward_integration(ctx, oracle):
    """cogsworth_b200.rewind.rewind_orbits = the batch form of hydro/rewind.py's _tohspans fan-out: every particle
    integrated from the snapshot time back to its own formation time (dt = -1 Myr, DOPRI853, final state)."""
    from cogsworth_b200.rewind import rewind_orbits
    pop = make_population(400, cfg_id=61, f_sn=0.0, horizon_gyr=0.3)
    b = pop.batch
    n = b.n_orbits
    rng = np.random.default_rng(5)
    t_form = 13000.0 - rng.uniform(5.0, 400.0, n)                       # Myr, ragged spans incl. non-integer ones
    pos, vel_kms = b.w0[:3], b.w0[3:] * K.KMS_PER_KPC_MYR
    for integ, o_int, tol in (("dopri853", oracle.DOP853, 1e-7), ("leapfrog", oracle.LEAPFROG, 1e-10),
                              ("dop853_dense", oracle.DOP853_DENSE, 1e-7)):
        p0, v0, ok = rewind_orbits(pos, vel_kms, 13000.0, t_form, -1.0, potential=pop.potential, integrator=integ, context=ctx)
        assert ok.all() and p0.shape == (n, 3)
        ref = oracle.integrate_batch(pop.potential, b.w0, np.full(n, 13000.0), t_form, np.full(n, -1.0), 1.0, 1,
                                     integrator=o_int)
        assert (ref["status"] == 0).all()
        d = np.maximum(np.abs(p0.T - ref["w_final"][:3]).max(axis=0) / np.linalg.norm(ref["w_final"][:3], axis=0),
                       np.abs(v0.T / K.KMS_PER_KPC_MYR - ref["w_final"][3:]).max(axis=0) / np.linalg.norm(ref["w_final"][3:], axis=0))
        assert np.quantile(d, 0.95) < tol, (integ, np.quantile(d, [0.5, 0.95, 1]))
    # node counts follow gala's backward grid: ceil(span) nodes above t2, plus t2 itself
    nn, _ = ctx.count_nodes(b.w0, np.full(n, 13000.0), t_form, np.full(n, -1.0), np.ones(n), np.ones(n, dtype=np.int32))
    assert np.array_equal(nn, np.ceil(13000.0 - t_form).astype(np.int32) + 1)
    # forward then backward returns to the start (leapfrog is time-reversible)
    fwd = ctx.integrate(b.w0, np.zeros(n), np.full(n, 200.0), np.ones(n), np.ones(n), np.zeros(n, dtype=np.int32),
                        integrator=_lib.LEAPFROG)["w_final"]            # gala's own grid: ends on node 199
    p1, v1, ok = rewind_orbits(fwd[:3], fwd[3:] * K.KMS_PER_KPC_MYR, 199.0, np.zeros(n), -1.0, potential=pop.potential,
                               integrator="leapfrog", context=ctx)
    assert np.abs(p1.T - b.w0[:3]).max() < 1e-10
    with pytest.raises(ValueError):
        rewind_orbits(pos, vel_kms, 100.0, np.full(n, 200.0), -1.0, potential=pop.potential, context=ctx)


def test_chunked_host_store_all_matches_device_mode(ctx):
    """Host-buffer store_all calls large enough to be cut into pipelined chunks (per-chunk event and trajectory
    column bases) return exactly the bits of one device-resident launch."""
    import torch
    pop = population_with_n_orbits(300_000, cfg_id=33, horizon_gyr=0.012)       # 13 nodes per orbit
    b = pop.batch
    n = b.n_orbits
    ctx.set_potential(pop.potential)
    host = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                         integrator=_lib.LEAPFROG, store_all=True)
    assert ctx.last_launch_count > 6                                           # several shards
    off = host["traj_off"]
    dev = torch.device("cuda:0")
    tt = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
    d = dict(w0=tt(b.w0), t1=tt(b.t1), t2=tt(b.t2), dt=tt(b.dt), to_myr=tt(b.to_myr), flags=tt(b.flags, torch.int32),
             ev_off=tt(b.ev_off, torch.int64), ev_time=tt(b.ev_time), ev_dv=tt(b.ev_dv))
    wf = torch.empty((6, n), dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    toff = tt(off, torch.int64)
    tp = torch.full((3, int(off[-1])), float("nan"), dtype=torch.float64, device=dev)
    tv = torch.full((3, int(off[-1])), float("nan"), dtype=torch.float64, device=dev)
    tm = torch.empty(int(off[-1]), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ctx.integrate_raw(n, d["w0"], d["t1"], d["t2"], d["dt"], d["to_myr"], d["flags"], d["ev_off"], d["ev_time"],
                      d["ev_dv"], wf, st, None, None, toff, tp, tv, tm, None, integrator=_lib.LEAPFROG,
                      mem_space=_lib.MEM_DEVICE)
    assert np.array_equal(tp.cpu().numpy(), host["traj_pos"])
    assert np.array_equal(tv.cpu().numpy(), host["traj_vel"])
    assert np.array_equal(tm.cpu().numpy(), host["traj_t"])
    assert np.array_equal(wf.cpu().numpy(), host["w_final"])


def test_store_all_full_size_properties(ctx, oracle):
    """configs[2] at FULL size on one B200: 10^6 orbits x 1000 stored steps = 48 GB of trajectories, device
    resident.  Size-independent properties checked where the data lives: no unwritten column, first sample = w0,
    last sample = w_final, the final-state kernel returns the same bits, and 200 000 randomly chosen samples are
    one oracle KDK step from their predecessors."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs ~55 GB of free device memory")
    pop = population_with_n_orbits(1_000_000, cfg_id=3, horizon_gyr=1.0, f_sn=0.0)
    b = pop.batch
    n = b.n_orbits
    ctx.set_potential(pop.potential)
    nn, _ = ctx.count_nodes(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, None, None, None)
    assert (nn == 1000).all()
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nn, out=off[1:])
    dev = torch.device("cuda:0")
    tt = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
    d = dict(w0=tt(b.w0), t1=tt(b.t1), t2=tt(b.t2), dt=tt(b.dt), to_myr=tt(b.to_myr), flags=tt(b.flags, torch.int32))
    toff = tt(off, torch.int64)
    total = int(off[-1])
    tp = torch.full((3, total), float("nan"), dtype=torch.float64, device=dev)
    tv = torch.full((3, total), float("nan"), dtype=torch.float64, device=dev)
    wf = torch.empty((6, n), dtype=torch.float64, device=dev)
    wf2 = torch.empty((6, n), dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ctx.integrate_raw(n, d["w0"], d["t1"], d["t2"], d["dt"], d["to_myr"], d["flags"], None, None, None, wf, st, None, None,
                      toff, tp, tv, None, stats, integrator=_lib.LEAPFROG, mem_space=_lib.MEM_DEVICE)
    k_ms = ctx.last_kernel_ms(0)
    assert int(stats[3]) == 999 * n and int((st != 0).sum()) == 0
    assert not bool(torch.isnan(tp).any()) and not bool(torch.isnan(tv).any())       # every column written
    first, last = toff[:-1], toff[1:] - 1
    assert bool((tp[:, first] == d["w0"][:3]).all()) and bool((tv[:, first] == d["w0"][3:]).all())
    assert bool((tp[:, last] == wf[:3]).all()) and bool((tv[:, last] == wf[3:]).all())
    ctx.integrate_raw(n, d["w0"], d["t1"], d["t2"], d["dt"], d["to_myr"], d["flags"], None, None, None, wf2, st, None, None,
                      None, None, None, None, stats, integrator=_lib.LEAPFROG, mem_space=_lib.MEM_DEVICE)
    assert bool((wf2 == wf).all())                                                   # final-state kernel: same bits
    # random local-step check against the oracle (columns that are not the last of their orbit)
    rng = np.random.default_rng(1)
    src = rng.integers(0, total, 200_000)
    src = src[(src % 1000) != 999]
    src_d = torch.from_numpy(src).to(dev)
    w_src = torch.cat([tp[:, src_d], tv[:, src_d]]).cpu().numpy()
    w_dst = torch.cat([tp[:, src_d + 1], tv[:, src_d + 1]]).cpu().numpy()
    m = len(src)
    o = oracle.integrate_batch(pop.potential, w_src, np.zeros(m), np.full(m, 1.5), np.ones(m), 1.0, 0, integrator=oracle.LEAPFROG)
    r = np.maximum(np.linalg.norm(w_dst[:3], axis=0), 1e-3)
    assert (np.abs(o["w_final"][:3] - w_dst[:3]).max(axis=0) / r).max() < 1e-9
    assert np.abs(o["w_final"][3:] - w_dst[3:]).max() < 1e-9
    assert 48.0 * total / (k_ms * 1e-3) / 1e9 > 2000.0                               # and it ran at HBM-class speed
    del tp, tv
    torch.cuda.empty_cache()


def test_host_store_all_streams_through_bounded_buffers(ctx, monkeypatch):
    """Host-buffer store_all calls are cut by trajectory columns and every stream lane re-uses its own device
    buffers shard after shard (device memory bounded by three shards, not by the batch): forcing hundreds of
    tiny shards must not change a bit, for ragged orbit lengths and with kicks."""
    pop = make_population(700, cfg_id=35)                                       # Wagg horizons: 10^2 .. 1.2x10^4 nodes
    b = pop.batch
    ctx.set_potential(pop.potential)
    args = (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
    ref = ctx.integrate(*args, integrator=_lib.LEAPFROG, store_all=True)
    n_ref = ctx.last_launch_count
    monkeypatch.setenv("CW_TRAJ_SHARD_COLS", "60000")
    cut = ctx.integrate(*args, integrator=_lib.LEAPFROG, store_all=True)
    assert ctx.last_launch_count > 20 * n_ref                                   # hundreds of shards
    for k in ("traj_pos", "traj_vel", "traj_t", "w_final", "status", "n_nodes", "ev_step"):
        assert np.array_equal(ref[k], cut[k]), k
    assert int(cut["stats"][3]) == int(ref["stats"][3])
    cut2 = ctx.integrate(*args, integrator=_lib.DOP853_DENSE, store_all=True)   # single shard by design, same API
    monkeypatch.delenv("CW_TRAJ_SHARD_COLS")
    ref2 = ctx.integrate(*args, integrator=_lib.DOP853_DENSE, store_all=True)
    assert np.array_equal(ref2["traj_pos"], cut2["traj_pos"])


def test_trajectory_slot_too_small_is_reported_not_overrun(ctx):
    """A caller-supplied traj_off whose slot is smaller than an orbit's node count: that orbit is rejected with
    CW_ORBIT_CAPACITY (-6) and every other orbit's slot holds exactly its trajectory (no overrun into a neighbour)."""
    pop = make_population(40, cfg_id=71, f_sn=0.0, horizon_gyr=0.05)            # 50 nodes each
    b = pop.batch
    n = b.n_orbits
    ctx.set_potential(pop.potential)
    good = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, integrator=_lib.LEAPFROG, store_all=True)
    lens = good["n_nodes"].astype(np.int64).copy()
    lens[3] -= 7                                                                 # too small
    lens[5] += 4                                                                 # larger than needed: fine
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tp = np.full((3, off[-1]), np.nan); tv = np.full((3, off[-1]), np.nan); tt = np.full(off[-1], np.nan)
    wf = np.empty((6, n)); st = np.empty(n, dtype=np.int32); nn = np.empty(n, dtype=np.int32)
    ctx.integrate_raw(n, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, None, None, None, wf, st, nn, None, off, tp, tv, tt, None,
                      integrator=_lib.LEAPFROG, mem_space=_lib.MEM_HOST)
    assert st[3] == -6 and (np.delete(st, 3) == 0).all()
    # (the columns of a rejected orbit, and the slack of an over-sized slot, are unspecified: host-buffer calls
    # copy the device buffer of the whole shard back)
    goff = good["traj_off"]
    for i in (0, 2, 4, 5, n - 1):
        m = good["n_nodes"][i]
        assert np.array_equal(tp[:, off[i]:off[i] + m], good["traj_pos"][:, goff[i]:goff[i] + m])
    assert np.array_equal(wf[:, 5], good["w_final"][:, 5]) and np.array_equal(nn, good["n_nodes"])


def test_steps_either_side_of_the_path(ctx, oracle):
    """birth-time circular velocities (pop.py:1001-1004) and the `escaped` mask (pop.py:874-895)
    from the pointwise kernels, in a static and in an evolving potential."""
    from test_host import fake_population
    from cogsworth_b200 import potential as P
    p = fake_population()
    p._cw_context = ctx
    ig = p.initial_galaxy
    q = np.stack([ig.x, ig.y, ig.z]).astype(float)
    for pot in (MilkyWayPotential("v2"), P.evolving_milky_way(np.linspace(0, 12000, 5), np.linspace(0.1, 1.0, 5))):
        p.galactic_potential = pot
        vc = p.initial_circular_velocity()
        t_birth = (float(p.max_ev_time) - np.asarray(ig.tau, dtype=float)) * 1000.0
        for i in range(len(p)):
            import dataclasses
            want = oracle.circular_velocity(dataclasses.replace(pot, t_eval=float(t_birth[i])), q[:, i].copy())
            assert abs(vc[i] - want * K.KMS_PER_KPC_MYR) < 1e-12 * vc[i]
    # the evolving galaxy is lighter at birth: smaller circular velocities than today's
    p.galactic_potential = MilkyWayPotential("v2")
    np.random.seed(3)
    p.perform_galactic_evolution(progress_bar=False)
    esc = p.escaped
    assert esc.shape == (len(p.final_pos),) and esc.dtype == bool
    fp, fv = np.asarray(p.final_pos, dtype=float), np.asarray(p.final_vel, dtype=float)
    for i in range(len(fp)):
        v_esc = np.sqrt(-2.0 * oracle.value(p.galactic_potential, fp[i].copy())) * K.KMS_PER_KPC_MYR
        assert esc[i] == (np.linalg.norm(fv[i]) >= v_esc)
    # classify._get_rel_speed (classify.py:117-155) on the same orbits: circular velocity from the kernel, cylindrical
    # components and the reference's combination on the host
    from cogsworth_b200.classify import get_rel_speed
    rel = get_rel_speed(p.orbits, p.galactic_potential, context=ctx)
    assert rel.shape == (len(fp),)
    for i in range(len(fp)):
        vc = oracle.circular_velocity(p.galactic_potential, fp[i].copy()) * K.KMS_PER_KPC_MYR
        phi = np.arctan2(fp[i, 1], fp[i, 0])
        v_R = fv[i, 0] * np.cos(phi) + fv[i, 1] * np.sin(phi)
        v_T = -fv[i, 0] * np.sin(phi) + fv[i, 1] * np.cos(phi)
        assert abs(rel[i] - np.sqrt((v_R - vc) ** 2 + v_T ** 2 + fv[i, 2] ** 2)) < 1e-9 * rel[i]


def test_population_in_an_evolving_potential(ctx):
    """evolving_potentials.ipynb cells 11 and 16: `Population(..., galactic_potential=evolving)` and swapping the
    potential of an evolved population, then `perform_galactic_evolution()` again."""
    from test_host import fake_population
    from cogsworth_b200 import potential as P
    p = fake_population()
    p._cw_context = ctx
    np.random.seed(5)
    p.perform_galactic_evolution(progress_bar=False)
    static = np.array(p.final_pos, dtype=float).copy()
    p.galactic_potential = P.evolving_milky_way(np.linspace(0, 12000, 5), np.linspace(0.1, 1.0, 5))
    p.perform_galactic_evolution(progress_bar=False)             # same binaries, same kicks (angles persisted)
    evolving = np.array(p.final_pos, dtype=float)
    assert evolving.shape == static.shape and np.isfinite(evolving).all()
    assert np.linalg.norm(evolving - static, axis=1).min() > 0.1     # a lighter young galaxy: different orbits
    p.integrator = None                                              # the reference's default, DOPRI853
    p.perform_galactic_evolution(progress_bar=False)
    dop = np.array(p.final_pos, dtype=float)
    assert np.isfinite(dop).all() and np.abs(dop - evolving).max() < 0.5   # leapfrog vs DOPRI853 of the same orbits
