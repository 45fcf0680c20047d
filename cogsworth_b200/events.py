"""Supernova event tables: host-side mirror of ``cogsworth.events.identify_events``
(reference ``src/cogsworth/events.py:6-72``).

The reference joins three pandas tables per call; here the join is done once on integer keys
with numpy (``bin_num * 4 + star``) so it scales to 10^7 systems, and the result is available
both as the reference's DataFrames (index = ``bin_num``; columns ``tphys, delta_vsys_x,
delta_vsys_y, delta_vsys_z, inc, phase``) and as the flat :class:`~cogsworth_b200.batch.EventTable`
the kernels consume.
"""
from __future__ import annotations

import numpy as np

from .batch import EventTable

__all__ = ["identify_events", "identify_event_tables", "draw_sn_angles"]

_COLS = ["tphys", "delta_vsys_x", "delta_vsys_y", "delta_vsys_z", "inc", "phase"]


def draw_sn_angles(initC, rng=None):
    """Draw and persist per-SN phase / inclination if absent (pop.py:1229-1234, events.py:21-26).

    Uses the global ``np.random`` state like the reference unless ``rng`` is given."""
    n = len(initC)
    for col in ("phase_sn_1", "phase_sn_2"):
        if col not in initC:
            initC[col] = np.random.uniform(0, 2 * np.pi, n) if rng is None else rng.uniform(0, 2 * np.pi, n)
    for col in ("inc_sn_1", "inc_sn_2"):
        if col not in initC:
            u01 = np.random.rand(n) if rng is None else rng.random(n)
            initC[col] = np.arccos(2 * u01 - 1.0)
    return initC


def _sn_rows(p):
    """All supernova rows joined with their kicks and angles, in bpp order.

    Returns a dict of equal-length arrays."""
    kick = p.kick_info
    bpp = p.bpp
    initC = p.initC if hasattr(p, "initC") else p.initial_binaries
    draw_sn_angles(initC)

    k_star = np.asarray(kick["star"].values)
    k_idx = np.flatnonzero(k_star > 0.0)
    k_key = np.asarray(kick["bin_num"].values)[k_idx].astype(np.int64) * 4 + k_star[k_idx].astype(np.int64)

    et = np.asarray(bpp["evol_type"].values)
    b_idx = np.flatnonzero((et == 15) | (et == 16))
    b_bin = np.asarray(bpp["bin_num"].values)[b_idx].astype(np.int64)
    b_star = np.where(et[b_idx] == 15, 1, 2).astype(np.int64)
    b_key = b_bin * 4 + b_star
    b_tphys = np.asarray(bpp["tphys"].values, dtype=np.float64)[b_idx]

    # inner join bpp SN rows with kick rows on (bin_num, star); keys are unique on the kick side
    if len(k_key) == len(b_key) and np.array_equal(k_key, b_key):
        # COSMIC writes both tables in the same (bin_num, time) order: row k of one IS row k of the other
        hit, krow = slice(None), k_idx
    else:
        order = None
        ks = k_key
        if len(ks) > 1 and np.any(ks[1:] < ks[:-1]):
            order = np.argsort(ks, kind="stable")
            ks = ks[order]
        pos = np.searchsorted(ks, b_key)
        pos_c = np.minimum(pos, max(len(ks) - 1, 0))
        hit = (pos < len(ks)) & (ks[pos_c] == b_key) if len(ks) else np.zeros(len(b_key), dtype=bool)
        krow = pos_c[hit] if order is None else order[pos_c[hit]]
        krow = k_idx[krow] if len(ks) else np.zeros(0, dtype=np.int64)

    def kcol(name):
        return np.asarray(kick[name].values, dtype=np.float64)[krow]

    bins = b_bin[hit]
    star = b_star[hit]
    # angles: row of initC for each bin_num (initC is indexed by bin_num)
    ic_bins = np.asarray(initC["bin_num"].values if "bin_num" in initC else initC.index.values).astype(np.int64)
    if len(ic_bins) and ic_bins[-1] - ic_bins[0] == len(ic_bins) - 1 and (len(ic_bins) < 2 or np.all(ic_bins[1:] > ic_bins[:-1])):
        ic_row = bins - ic_bins[0]                    # consecutive bin_nums: the row is the offset
    elif len(ic_bins) > 1 and np.any(ic_bins[1:] < ic_bins[:-1]):
        ic_order = np.argsort(ic_bins, kind="stable")
        ic_row = ic_order[np.searchsorted(ic_bins[ic_order], bins)]
    else:
        ic_row = np.searchsorted(ic_bins, bins)
    inc1, inc2 = np.asarray(initC["inc_sn_1"].values), np.asarray(initC["inc_sn_2"].values)
    ph1, ph2 = np.asarray(initC["phase_sn_1"].values), np.asarray(initC["phase_sn_2"].values)
    K = len(bins)
    dv1, dv2 = np.empty((K, 3)), np.empty((K, 3))
    for c, ax in enumerate("xyz"):
        dv1[:, c] = kcol(f"delta_vsys{ax}_1")
        dv2[:, c] = kcol(f"delta_vsys{ax}_2")
    first = star == 1
    return dict(
        bin_num=bins, star=star, tphys=b_tphys[hit], disrupted=kcol("disrupted"), dv1=dv1, dv2=dv2,
        inc=np.where(first, inc1[ic_row], inc2[ic_row]).astype(np.float64, copy=False),
        phase=np.where(first, ph1[ic_row], ph2[ic_row]).astype(np.float64, copy=False))


def _select(rows, mask, dv):
    return dict(bin_num=rows["bin_num"][mask], tphys=rows["tphys"][mask], dv=dv[mask],
                inc=rows["inc"][mask], phase=rows["phase"][mask])


def _primary_secondary(p):
    rows = _sn_rows(p)
    # primaries: delta_vsys*_1 at every supernova                              (events.py:47-53)
    prim = dict(bin_num=rows["bin_num"], tphys=rows["tphys"], dv=rows["dv1"], inc=rows["inc"], phase=rows["phase"])
    # secondaries: *_2, except *_1 while the binary is still bound             (events.py:56-67)
    dv_sec = np.where((rows["disrupted"] == 0)[:, None], rows["dv1"], rows["dv2"])
    # ... and only for disrupted binaries, in bin_nums[disrupted] order         (events.py:70)
    bin_nums = np.asarray(p.bin_nums).astype(np.int64)
    dis_bins = bin_nums[np.asarray(p.disrupted, dtype=bool)]
    sec_idx = np.zeros(0, dtype=np.int64)
    if len(dis_bins) and len(rows["bin_num"]):
        sb = rows["bin_num"]
        o = None
        if len(sb) > 1 and np.any(sb[1:] < sb[:-1]):        # bpp is normally grouped by bin_num already
            o = np.argsort(sb, kind="stable")
            sb = sb[o]
        lo, hi = np.searchsorted(sb, dis_bins, "left"), np.searchsorted(sb, dis_bins, "right")
        lens = hi - lo
        ends = np.cumsum(lens)
        sec_idx = np.repeat(lo - (ends - lens), lens) + np.arange(int(ends[-1]) if len(ends) else 0)
        if o is not None:
            sec_idx = o[sec_idx]
    m = np.zeros(len(rows["bin_num"]), dtype=bool)
    sec = _select(rows, sec_idx, dv_sec) if len(sec_idx) else _select(rows, m, dv_sec)
    return prim, sec, bin_nums


def _frame(t):
    import pandas as pd
    df = pd.DataFrame({"tphys": t["tphys"], "delta_vsys_x": t["dv"][:, 0], "delta_vsys_y": t["dv"][:, 1],
                       "delta_vsys_z": t["dv"][:, 2], "inc": t["inc"], "phase": t["phase"]}, columns=_COLS)
    df.index = t["bin_num"]
    return df


def identify_events(p):
    """Drop-in for ``cogsworth.events.identify_events(p)``: (primary_event_rows, secondary_event_rows)."""
    prim, sec, _ = _primary_secondary(p)
    return _frame(prim), _frame(sec)


def _table(t, bin_nums):
    """EventTable whose owners are positions in ``bin_nums``, grouped by owner, time order kept."""
    if not len(t["bin_num"]):
        return EventTable.empty()
    if len(bin_nums) > 1 and np.any(bin_nums[1:] < bin_nums[:-1]):
        order_b = np.argsort(bin_nums, kind="stable")
        owner = order_b[np.searchsorted(bin_nums[order_b], t["bin_num"])]
    elif bin_nums[-1] - bin_nums[0] == len(bin_nums) - 1 and (len(bin_nums) < 2 or np.all(bin_nums[1:] > bin_nums[:-1])):
        owner = t["bin_num"] - bin_nums[0]    # consecutive bin_nums
    else:                                     # the population is in bin_num order (the usual case)
        owner = np.searchsorted(bin_nums, t["bin_num"])
    owner = owner.astype(np.int64, copy=False)
    if len(owner) > 1 and np.any(owner[1:] < owner[:-1]):
        o = np.argsort(owner, kind="stable")  # stable: keeps the chronological bpp order per system
        return EventTable(owner[o], t["tphys"][o], t["dv"][o], t["phase"][o], t["inc"][o])
    return EventTable(owner, t["tphys"], t["dv"], t["phase"], t["inc"])


class _Range:
    """The rows of a population whose bin_num lies in one contiguous range: what the joins above read, sliced."""

    def __init__(self, kick_info, bpp, initC, bin_nums, disrupted):
        self.kick_info, self.bpp, self.initC, self.bin_nums, self.disrupted = kick_info, bpp, initC, bin_nums, disrupted


#: populations at least this large are joined range by range on a thread pool (numpy releases the GIL in every pass)
PARALLEL_MIN_BINARIES = 1_000_000


def _tables_by_range(p, n_threads):
    """identify_event_tables for a population in bin_num order whose tables are grouped by bin_num (what COSMIC
    writes): the population is cut into contiguous bin_num ranges, every range is joined on its own by the functions
    above, concurrently, and the tables are concatenated -- the same rows in the same order as the one-pass join.
    Returns None when the orderings do not allow it."""
    from concurrent.futures import ThreadPoolExecutor
    bin_nums = np.asarray(p.bin_nums).astype(np.int64)
    n = len(bin_nums)
    kick, bpp = p.kick_info, p.bpp
    initC = p.initC if hasattr(p, "initC") else p.initial_binaries
    kb = np.asarray(kick["bin_num"].values)
    bb = np.asarray(bpp["bin_num"].values)
    ib = np.asarray(initC["bin_num"].values if "bin_num" in initC else initC.index.values)
    disrupted = np.asarray(p.disrupted, dtype=bool)
    cuts = [(n * k) // n_threads for k in range(n_threads + 1)]
    edges = bin_nums[[min(c, n - 1) for c in cuts[1:-1]]]          # first bin_num of ranges 1 .. T-1
    ka = [0] + list(np.searchsorted(kb, edges, "left")) + [len(kb)]
    ba = [0] + list(np.searchsorted(bb, edges, "left")) + [len(bb)]
    ia = [0] + list(np.searchsorted(ib, edges, "left")) + [len(ib)]

    def sorted_slice(a, lo, hi, first, last):
        v = a[lo:hi]
        return len(v) == 0 or (bool(np.all(v[1:] >= v[:-1])) and v[0] >= first and (last is None or v[-1] < last))

    def work(k):
        ca, cb = cuts[k], cuts[k + 1]
        if cb <= ca:
            return EventTable.empty(), EventTable.empty()
        first = bin_nums[ca]
        last = bin_nums[cb] if cb < n else None
        bn = bin_nums[ca:cb]
        # the cut is only valid if every table really is ordered by bin_num (searchsorted above assumed it)
        if not (bool(np.all(bn[1:] > bn[:-1])) and (last is None or bn[-1] < last) and
                sorted_slice(kb, ka[k], ka[k + 1], first, last) and sorted_slice(bb, ba[k], ba[k + 1], first, last) and
                sorted_slice(ib, ia[k], ia[k + 1], first, last)):
            return None
        r = _Range(kick.iloc[ka[k]:ka[k + 1]], bpp.iloc[ba[k]:ba[k + 1]], initC.iloc[ia[k]:ia[k + 1]], bn, disrupted[ca:cb])
        prim, sec, _ = _primary_secondary(r)
        tp, ts = _table(prim, bn), _table(sec, bn)
        for t in (tp, ts):
            if len(t):
                t.owner += ca                                    # positions in the whole population
        return tp, ts

    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        parts = list(ex.map(work, range(n_threads)))
    if any(part is None for part in parts):
        return None
    out = []
    for j in (0, 1):
        ts = [part[j] for part in parts if len(part[j])]
        if not ts:
            out.append(EventTable.empty())
        elif len(ts) == 1:
            out.append(ts[0])
        else:
            out.append(EventTable(np.concatenate([t.owner for t in ts]), np.concatenate([t.tphys for t in ts]),
                                  np.concatenate([t.dv for t in ts]), np.concatenate([t.phase for t in ts]),
                                  np.concatenate([t.inc for t in ts])))
    return out[0], out[1]


def identify_event_tables(p, n_threads=None):
    """Same content as :func:`identify_events` as flat :class:`EventTable` s (owners = positions in
    ``p.bin_nums``), ready for :func:`cogsworth_b200.batch.build_orbit_batch`.  Large populations in bin_num order are
    joined range by range on ``n_threads`` host threads (default: the cores this process may use, at most 16)."""
    n = len(p.bin_nums)
    if n_threads is None:
        import os
        try:
            n_threads = min(len(os.sched_getaffinity(0)), 16)
        except AttributeError:
            n_threads = min(os.cpu_count() or 1, 16)
    if n >= PARALLEL_MIN_BINARIES and n_threads > 1:
        draw_sn_angles(p.initC if hasattr(p, "initC") else p.initial_binaries)       # once, before the ranges read them
        res = _tables_by_range(p, n_threads)
        if res is not None:
            return res
    prim, sec, bin_nums = _primary_secondary(p)
    return _table(prim, bin_nums), _table(sec, bin_nums)


def event_table_from_frame(df, bin_nums):
    """Convert a reference-style events DataFrame (index = bin_num) to an EventTable."""
    t = dict(bin_num=np.asarray(df.index.values).astype(np.int64),
             tphys=np.asarray(df["tphys"].values, dtype=np.float64),
             dv=np.stack([np.asarray(df[c].values, dtype=np.float64) for c in _COLS[1:4]], axis=1)
             if len(df) else np.zeros((0, 3)),
             inc=np.asarray(df["inc"].values, dtype=np.float64),
             phase=np.asarray(df["phase"].values, dtype=np.float64))
    return _table(t, np.asarray(bin_nums).astype(np.int64))
