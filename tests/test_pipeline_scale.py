"""The table-driven stage at scale: 10^4 binaries of bpp / kick_info / initial_binaries
DataFrames through identify_event_tables -> build_orbit_batch (-> GPU), against a LITERAL per-binary restatement of what
the reference does with the same tables:

  * `literal_identify_events`   pandas merges, row for row what events.py:6-72 computes (written independently of
                                cogsworth_b200.events, which joins integer keys with numpy);
  * `literal_orbit_args`        the generator of pop.py:1156-1193: primaries in bin_num order, then the disrupted
                                secondaries, events looked up per binary with `.loc`;
  * `literal_batch`             the unit arithmetic integrate_orbit_with_events applies to each of those argument
                                tuples (kicks.py:97,118-122,134,160-172), one orbit at a time in Python floats.

The CPU test demands the vectorised pipeline's arrays to be BIT-IDENTICAL to the literal ones; the GPU test runs
B200Population.perform_galactic_evolution() on the tables and compares with the oracle integrating the literal arrays.
"""
import math

import numpy as np
import pandas as pd
import pytest

from cogsworth_b200.batch import build_orbit_batch
from cogsworth_b200.events import identify_event_tables, identify_events
from cogsworth_b200.pop import B200Population
from cogsworth_b200.synthetic import make_population, population_tables

N_BINARIES = 10_000


def literal_identify_events(p):
    """(primary rows, secondary rows) as DataFrames indexed by bin_num -- events.py:6-72 restated with plain merges."""
    k = p.kick_info
    k = k[k["star"] > 0][["bin_num", "star", "disrupted"] + [f"delta_vsys{a}_{s}" for s in (1, 2) for a in "xyz"]]
    b = p.bpp
    b = b[(b["evol_type"] == 15) | (b["evol_type"] == 16)][["bin_num", "tphys", "evol_type"]].copy()
    b["star"] = np.where(b["evol_type"] == 15, 1.0, 2.0)
    ic = p.initial_binaries[["bin_num", "inc_sn_1", "inc_sn_2", "phase_sn_1", "phase_sn_2"]]
    j = b.merge(k, on=["bin_num", "star"], how="inner", sort=False).merge(ic, on="bin_num", how="left", sort=False)
    j["inc"] = np.where(j["star"] == 1, j["inc_sn_1"], j["inc_sn_2"])
    j["phase"] = np.where(j["star"] == 1, j["phase_sn_1"], j["phase_sn_2"])
    j.index = j["bin_num"].values
    prim = pd.DataFrame({"tphys": j["tphys"], "delta_vsys_x": j["delta_vsysx_1"], "delta_vsys_y": j["delta_vsysy_1"],
                         "delta_vsys_z": j["delta_vsysz_1"], "inc": j["inc"], "phase": j["phase"]})
    bound = (j["disrupted"] == 0).values
    sec = pd.DataFrame({"tphys": j["tphys"],
                        "delta_vsys_x": np.where(bound, j["delta_vsysx_1"], j["delta_vsysx_2"]),
                        "delta_vsys_y": np.where(bound, j["delta_vsysy_1"], j["delta_vsysy_2"]),
                        "delta_vsys_z": np.where(bound, j["delta_vsysz_1"], j["delta_vsysz_2"]),
                        "inc": j["inc"], "phase": j["phase"]}, index=j.index)
    sec = sec.loc[p.bin_nums[p.disrupted]] if p.disrupted.any() else sec.iloc[:0]
    return prim, sec


def literal_orbit_args(p, prim, sec):
    """pop.py:1156-1193: (w0 [kpc, km/s], t1 [Gyr], dt [Myr], events | None) per orbit, primaries then secondaries."""
    g = p.initial_galaxy
    w0s = np.stack([g.x, g.y, g.z, g.v_x, g.v_y, g.v_z])
    has = set(prim.index.values.tolist())
    for i in range(len(p)):
        bn = p.bin_nums[i]
        yield w0s[:, i], p.max_ev_time - g.tau[i], p.timestep_size, (prim.loc[[bn]] if bn in has else None)
    for i in np.arange(len(p))[p.disrupted]:
        yield w0s[:, i], p.max_ev_time - g.tau[i], p.timestep_size, sec.loc[[p.bin_nums[i]]]


SEC_PER_MYR, SEC_PER_GYR, KMS = 3.15576e13, 3.15576e16, 977.7922216807891


def literal_batch(args, t2_gyr):
    """kicks.py:97-172 for every argument tuple: the plain cw_batch columns, computed orbit by orbit."""
    w0, t1, t2, dt, tm, fl, off, et, ed = [], [], [], [], [], [], [0], [], []
    for w, t1_gyr, dt_myr, ev in args:
        t1_gyr, dt_myr = float(t1_gyr), float(dt_myr)
        span = t2_gyr - t1_gyr
        clipped = span < dt_myr * (SEC_PER_MYR / SEC_PER_GYR)        # dt = min(dt, t2 - t1), kicks.py:97
        w0.append(np.concatenate([w[:3], w[3:] * (1.0 / KMS)]))
        if ev is None:                                               # gala builds the grid itself, in Myr
            t1.append(t1_gyr * 1000.0); t2.append(t2_gyr * 1000.0)
            dt.append(span * 1000.0 if clipped else dt_myr); tm.append(1.0); fl.append(0)
        else:                                                        # seconds, t2 appended (kicks.py:118-122)
            t1.append(t1_gyr * SEC_PER_GYR); t2.append(t2_gyr * SEC_PER_GYR)
            dt.append(span * SEC_PER_GYR if clipped else dt_myr * SEC_PER_MYR); tm.append(1.0 / SEC_PER_MYR); fl.append(1)
            for _, r in ev.iterrows():
                et.append((t1_gyr + r["tphys"] * (SEC_PER_MYR / SEC_PER_GYR)) * SEC_PER_GYR)      # kicks.py:134
                vx, vy, vz, th, ph = r["delta_vsys_x"], r["delta_vsys_y"], r["delta_vsys_z"], r["phase"], r["inc"]
                st, ct, sp, cp = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)                   # kicks.py:42-46
                ed.append([(vx * ct - vy * st * cp + vz * st * sp) * (1.0 / KMS),
                           (vx * st + vy * ct * cp - vz * ct * sp) * (1.0 / KMS),
                           (vy * sp + vz * cp) * (1.0 / KMS)])
        off.append(len(et))
    return dict(w0=np.ascontiguousarray(np.array(w0).T), t1=np.array(t1), t2=np.array(t2), dt=np.array(dt),
                to_myr=np.array(tm), flags=np.array(fl, dtype=np.int32), ev_off=np.array(off, dtype=np.int64),
                ev_time=np.array(et), ev_dv=np.array(ed).reshape(-1, 3))


@pytest.fixture(scope="module")
def tables():
    # horizons from 0.2 Myr: a few systems are younger than one time-step (dt is clipped to their span, kicks.py:97);
    # bin_nums do not start at zero
    pop = make_population(N_BINARIES, cfg_id=21, horizon_gyr=(0.0002, 0.35), f_sn=0.7)
    galaxy, bpp, kick_info, initial_binaries = population_tables(pop, bin_num0=5000)
    P = B200Population(galaxy, bpp, kick_info, initial_binaries, galactic_potential=pop.potential, max_ev_time=12.0,
                       timestep_size=1.0, store_entire_orbits=False, integrator="leapfrog")
    prim, sec = literal_identify_events(P)
    lit = literal_batch(literal_orbit_args(P, prim, sec), 12.0)
    return pop, P, prim, sec, lit


def test_identify_events_matches_the_literal_merges(tables):
    pop, P, prim, sec, _ = tables
    p2, s2 = identify_events(P)
    for a, b in ((prim, p2), (sec, s2)):
        assert len(a) == len(b) > 3000
        assert np.array_equal(a.index.values, b.index.values)
        for col in ("tphys", "delta_vsys_x", "delta_vsys_y", "delta_vsys_z", "inc", "phase"):
            assert np.array_equal(a[col].values.astype(float), b[col].values.astype(float)), col


def test_vectorised_batch_is_bit_identical_to_the_per_binary_loop(tables):
    pop, P, _, _, lit = tables
    pe, se = identify_event_tables(P)
    g = P.initial_galaxy
    b = build_orbit_batch(np.stack([g.x, g.y, g.z]), np.stack([g.v_x, g.v_y, g.v_z]), g.tau, 12.0, 1.0, P.disrupted, pe, se)
    assert b.compact is not None and b.n_orbits == len(P) + int(P.disrupted.sum()) == lit["w0"].shape[1]
    assert (lit["flags"] == 0).sum() > 1000 and (lit["flags"] == 1).sum() > 5000      # both grid branches present
    assert 5 < (lit["dt"] * lit["to_myr"] < 1.0).sum() < 200 and len(b.compact["classes"]) > 7    # clipped time-steps
    for k in ("w0", "t1", "t2", "dt", "to_myr", "flags", "ev_off", "ev_time", "ev_dv"):
        assert np.array_equal(getattr(b, k), lit[k]), k
    # the compact form sends the same numbers in fewer bytes
    a = b.abi_arrays()
    sent = sum(v.nbytes for v in a.values() if isinstance(v, np.ndarray))
    plain = sum(lit[k].nbytes for k in lit)
    assert sent < 0.7 * plain


def test_range_parallel_join_equals_the_one_pass_join(tables, monkeypatch):
    """Large populations are joined range by range on host threads (events._tables_by_range): the same rows in the same
    order as the one-pass join, whatever the number of ranges; tables that are not grouped by bin_num fall back."""
    from cogsworth_b200 import events
    _, P, _, _, _ = tables
    one = events.identify_event_tables(P, n_threads=1)
    monkeypatch.setattr(events, "PARALLEL_MIN_BINARIES", 100)
    for T in (2, 3, 7, 16):
        assert events._tables_by_range(P, T) is not None
        par = events.identify_event_tables(P, n_threads=T)
        for a, b in zip(one, par):
            assert len(a) == len(b) > 3000
            for f in ("owner", "tphys", "dv", "phase", "inc"):
                assert np.array_equal(getattr(a, f), getattr(b, f)), (T, f)
    # kick_info in another order: the range cut is refused and the sorting join takes over, same tables
    Q = B200Population(P.initial_galaxy, P.bpp, P.kick_info.iloc[::-1], P.initial_binaries, galactic_potential=None,
                       max_ev_time=12.0, timestep_size=1.0, store_entire_orbits=False, integrator="leapfrog")
    Q.disrupted
    assert events._tables_by_range(Q, 4) is None
    for a, b in zip(one, events.identify_event_tables(Q, n_threads=4)):
        for f in ("owner", "tphys", "dv", "phase", "inc"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.gpu
def test_population_stage_on_gpu_matches_oracle_on_literal_arrays(tables, ctx, oracle):
    from _cases import assert_parity_outside_the_sensitive_set, rel_diff, sensitive_orbits
    from cogsworth_b200.batch import OrbitBatch
    pop, P, _, _, lit = tables
    P._cw_context = ctx
    P.perform_galactic_evolution(progress_bar=False)
    info = P._last_integration_info
    n, nd = len(P), int(P.disrupted.sum())
    assert len(P.orbits) == n + nd and P.final_pos.shape == (n + nd, 3)
    lb = OrbitBatch(**lit)
    ref = oracle.integrate_batch(pop.potential, lb.w0, lb.t1, lb.t2, lb.dt, lb.to_myr, lb.flags, lb.ev_off, lb.ev_time,
                                 lb.ev_dv, integrator=oracle.LEAPFROG)
    assert np.array_equal(info["ev_step"], ref["ev_step"]), "kick nodes must be bit-exact"
    assert np.array_equal(info["n_nodes"], ref["n_nodes"]) and (info["status"] == 0).all()
    w = np.concatenate([np.asarray(P.final_pos).T, np.asarray(P.final_vel).T / KMS])
    d = rel_diff(w, ref["w_final"])
    sens, _, _ = sensitive_orbits(oracle, pop.potential, lb, oracle.LEAPFROG, ref=ref)
    rep = assert_parity_outside_the_sensitive_set(d, sens, 1e-10, 0.03, "table pipeline")
    assert rep["median"] < 1e-13
    # Q7: primaries first, then the disrupted secondaries in bin_nums[disrupted] order; a secondary's orbit differs
    # from its system's only after the disrupting kick
    sec_of = np.flatnonzero(P.disrupted)
    assert np.array_equal(P.final_secondary_pos[~P.disrupted], P.final_primary_pos[~P.disrupted])
    assert not np.array_equal(P.final_pos[n:], P.final_pos[sec_of])
