"""Flattening a population into the structure-of-arrays batch the kernels consume.

This replaces the per-binary Python generator ``Population._iter_orbit_args`` (reference
``src/cogsworth/pop.py:1156-1193``) -- which does a pandas ``.loc`` per binary -- by one
vectorised pass, and reproduces the unit arithmetic astropy performs on the way into
``integrate_orbit_with_events`` (``src/cogsworth/kicks.py:97,118-122,134,160-172``) float for
float, because the kick node index is defined by FP64 comparisons of those values:

* kicked orbits (``events`` is not None): the grid is built in SECONDS from ``t1`` [Gyr],
  ``t2`` [Gyr], ``dt`` [Myr] and ``t2`` is appended as the last node; an event at ``tphys`` [Myr]
  is compared as ``(t1_Gyr + tphys * 1e-3) * 3.15576e16`` seconds;
* un-kicked orbits (``events is None``): gala builds the grid itself in Myr and does not
  append ``t2`` (reference quirk Q1).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import constants as K

GRID_APPEND_T2 = 1


@dataclass
class EventTable:
    """Supernova events of a set of systems, sorted by (system position, time).

    ``owner[k]`` is the position (0..n-1) of the system the event belongs to in the population
    order; ``tphys`` is in Myr after birth, ``dv`` the BSE-frame systemic velocity change in km/s
    and ``phase`` / ``inc`` the angles of ``get_kick_differential``.
    """
    owner: np.ndarray
    tphys: np.ndarray
    dv: np.ndarray      # (K, 3) km/s
    phase: np.ndarray
    inc: np.ndarray

    def __len__(self):
        return len(self.owner)

    @staticmethod
    def empty() -> "EventTable":
        z = np.zeros(0)
        return EventTable(np.zeros(0, dtype=np.int64), z, np.zeros((0, 3)), z, z)

    def csr(self, n: int):
        """Offsets (n+1,) of each owner's rows; rows must be grouped by owner in ascending order."""
        if len(self.owner) and np.any(np.diff(self.owner) < 0):
            raise ValueError("EventTable rows must be grouped by owner (ascending)")
        counts = np.bincount(self.owner, minlength=n).astype(np.int64)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return off


def rotate_kicks_threaded(dv_kms, phase, inc, scale, out):
    """``rotate_kicks(...) * scale`` on the library's host threads (``cw_host_rotate_kicks``): the rotation is four
    libm calls per event, most of the batch builder's time at 10^7 events.  Returns False when the library is not
    built (numpy is used then).  libm's last bit may differ from numpy's: the result is AN input of the integration,
    identical for whoever integrates the batch."""
    try:
        from . import _lib
        L = _lib.load()
    except Exception:
        return False
    dv_kms, phase, inc = (np.ascontiguousarray(a, dtype=np.float64) for a in (dv_kms, phase, inc))
    L.cw_host_rotate_kicks(len(phase), _lib.ptr(dv_kms), _lib.ptr(phase), _lib.ptr(inc), float(scale), _lib.ptr(out))
    return True


def rotate_kicks(dv_kms: np.ndarray, phase: np.ndarray, inc: np.ndarray, out=None) -> np.ndarray:
    """Vectorised ``get_kick_differential`` (kicks.py:42-46): BSE (vx,vy,vz) -> Galactocentric, km/s."""
    vx, vy, vz = dv_kms[:, 0], dv_kms[:, 1], dv_kms[:, 2]
    st, ct, sp, cp = np.sin(phase), np.cos(phase), np.sin(inc), np.cos(inc)
    if out is None:
        out = np.empty((len(phase), 3))
    out[:, 0] = vx * ct - vy * st * cp + vz * st * sp
    out[:, 1] = vx * st + vy * ct * cp - vz * ct * sp
    out[:, 2] = vy * sp + vz * cp
    return out


class OrbitBatch:
    """Everything ``cw_integrate`` needs for a set of orbits (see include/cogsworth_b200.h).

    Two representations of the same orbits:

    * the PLAIN layout -- ``w0`` (6, n) kpc & kpc/Myr, ``t1``, ``t2``, ``dt``, ``to_myr`` (n,) in the orbit's grid
      unit, ``flags`` (n,) int32, events CSR ``ev_off`` (n+1,) int64 / ``ev_time`` (K,) / ``ev_dv`` (K, 3);
    * the COMPACT form ``build_orbit_batch`` produces (``cw_batch`` ABI 2): the initial conditions and birth times
      of the ``n_w0`` systems once, ``w0_index`` naming the system every disrupted secondary starts from
      (pop.py:1186-1193), one byte of ``grid_class`` per orbit into a table of (t2, dt, to_myr, flags) and 32-bit
      event offsets -- 78 instead of 134 bytes per orbit across the host link for a kicked population.

    The plain arrays of a compact batch are expanded on first access (tests, the retry loop, the CPU oracle);
    assigning to one of them drops the compact form.
    """

    _PLAIN = ("w0", "t1", "t2", "dt", "to_myr", "flags", "ev_off")

    def __init__(self, w0=None, t1=None, t2=None, dt=None, to_myr=None, flags=None, ev_off=None, ev_time=None,
                 ev_dv=None, n_primary=0, secondary_of=None, meta=None, compact=None):
        d = self.__dict__
        d["_plain"] = dict(w0=w0, t1=t1, t2=t2, dt=dt, to_myr=to_myr, flags=flags, ev_off=ev_off)
        d["ev_time"], d["ev_dv"] = ev_time, ev_dv
        d["n_primary"] = n_primary          # the first n_primary orbits are bound systems / primaries (Q7)
        d["secondary_of"] = secondary_of    # population index of each disrupted secondary
        d["meta"] = {} if meta is None else meta
        d["compact"] = compact              # dict(w0, t1, n_w0, w0_index, grid_class, classes, ev_off32) or None

    # -- plain arrays, expanded from the compact form on demand ---------------------------------------
    def __getattr__(self, name):
        if name in OrbitBatch._PLAIN:
            v = self.__dict__["_plain"][name]
            if v is None and self.__dict__["compact"] is not None:
                v = self._expand(name)
                self.__dict__["_plain"][name] = v
            return v
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in OrbitBatch._PLAIN:
            if self.__dict__["compact"] is not None:        # the compact form no longer describes the batch
                for k in OrbitBatch._PLAIN:
                    getattr(self, k)
                self.__dict__["compact"] = None
            self.__dict__["_plain"][name] = value
        else:
            self.__dict__[name] = value

    def _expand(self, name):
        c = self.compact
        n_w0, idx = c["n_w0"], c["w0_index"]
        if name == "w0":
            w0 = np.empty((6, n_w0 + len(idx)))
            w0[:, :n_w0] = c["w0"]
            w0[:, n_w0:] = c["w0"][:, idx]
            return w0
        if name == "t1":
            return np.concatenate([c["t1"], c["t1"][idx]])
        if name == "ev_off":
            return None if c["ev_off32"] is None else c["ev_off32"].astype(np.int64)
        col = {"t2": "t2", "dt": "dt", "to_myr": "to_myr", "flags": "grid_flags"}[name]
        return np.ascontiguousarray(c["classes"][col][c["grid_class"]])

    @property
    def n_orbits(self) -> int:
        c = self.compact
        return c["n_w0"] + len(c["w0_index"]) if c is not None else self.w0.shape[1]

    @property
    def n_events(self) -> int:
        return 0 if self.ev_time is None else int(self.ev_time.shape[0])

    def abi_arrays(self) -> dict:
        """The arrays of one ``cw_integrate`` call, compact when the batch has that form."""
        c = self.compact
        if c is None:
            return dict(w0=self.w0, t1=self.t1, t2=self.t2, dt=self.dt, to_myr=self.to_myr, flags=self.flags,
                        ev_off=self.ev_off if self.n_events else None, ev_time=self.ev_time, ev_dv=self.ev_dv)
        K = self.n_events
        return dict(w0=c["w0"], t1=c["t1"], n_w0=c["n_w0"], w0_index=c["w0_index"] if len(c["w0_index"]) else None,
                    grid_class=c["grid_class"], classes=c["classes"], ev_off32=c["ev_off32"] if K else None,
                    ev_time=self.ev_time if K else None, ev_dv=self.ev_dv if K else None)

    def head(self, n_orbits: int) -> "OrbitBatch":
        """The first ``n_orbits`` orbits; keeps the compact form when only secondaries are dropped."""
        c = self.compact
        if c is None or n_orbits < c["n_w0"]:
            out = self.take(np.arange(n_orbits))
            out.n_primary = min(self.n_primary, n_orbits)
            return out
        k = n_orbits - c["n_w0"]
        K = int(c["ev_off32"][n_orbits]) if c["ev_off32"] is not None else 0
        c2 = dict(c, w0_index=c["w0_index"][:k], grid_class=c["grid_class"][:n_orbits],
                  ev_off32=None if c["ev_off32"] is None else c["ev_off32"][:n_orbits + 1])
        return OrbitBatch(ev_time=None if self.ev_time is None else self.ev_time[:K],
                          ev_dv=None if self.ev_dv is None else self.ev_dv[:K], n_primary=self.n_primary,
                          secondary_of=None if self.secondary_of is None else self.secondary_of[:k], compact=c2)

    def take(self, idx) -> "OrbitBatch":
        """Sub-batch in the plain layout (used by the retry loop for failed orbits)."""
        idx = np.asarray(idx, dtype=np.int64)
        ev_off = ev_time = ev_dv = None
        if self.ev_off is not None and self.n_events:
            counts = (self.ev_off[1:] - self.ev_off[:-1])[idx]
            ev_off = np.zeros(len(idx) + 1, dtype=np.int64)
            np.cumsum(counts, out=ev_off[1:])
            rows = np.repeat(self.ev_off[:-1][idx] - ev_off[:-1], counts) + np.arange(ev_off[-1])
            ev_time, ev_dv = self.ev_time[rows], self.ev_dv[rows]
            if ev_off[-1] == 0:
                ev_off = ev_time = ev_dv = None
        return OrbitBatch(np.ascontiguousarray(self.w0[:, idx]), self.t1[idx], self.t2[idx], self.dt[idx],
                          self.to_myr[idx], self.flags[idx], ev_off, ev_time, ev_dv)


def _grid_columns(t1_gyr, t2_gyr, dt_myr, kicked):
    """Per-orbit (t1, t2, dt, to_myr, flags) following the two reference branches."""
    t1_gyr = np.asarray(t1_gyr, dtype=np.float64)
    n = t1_gyr.shape[0]
    t2_gyr = np.broadcast_to(np.asarray(t2_gyr, dtype=np.float64), (n,))
    dt_myr = np.broadcast_to(np.asarray(dt_myr, dtype=np.float64), (n,))
    span_gyr = t2_gyr - t1_gyr
    # dt = min(dt, t2 - t1) -> (t2 - t1) < dt, astropy converts dt to the left operand's unit
    # (Gyr); the winner keeps its own unit                                          (kicks.py:97)
    clip = span_gyr < dt_myr * K.GYR_PER_MYR
    t1 = np.where(kicked, t1_gyr * K.SEC_PER_GYR, t1_gyr * K.MYR_PER_GYR)
    t2 = np.where(kicked, t2_gyr * K.SEC_PER_GYR, t2_gyr * K.MYR_PER_GYR)
    dt_s = np.where(clip, span_gyr * K.SEC_PER_GYR, dt_myr * K.SEC_PER_MYR)
    dt_m = np.where(clip, span_gyr * K.MYR_PER_GYR, dt_myr)
    dt = np.where(kicked, dt_s, dt_m)
    to_myr = np.where(kicked, K.MYR_PER_SEC, 1.0)
    flags = np.where(kicked, GRID_APPEND_T2, 0).astype(np.int32)
    return t1, t2, dt, to_myr, flags


def _grid_classes(t1_gyr, t2_gyr, dt_myr, kicked):
    """The grid columns of ``_grid_columns`` as (grid_class uint8 (n,), classes table) -- ``cw_grid_class`` of
    include/cogsworth_b200.h -- or ``None`` when more than 256 classes would be needed.  Same floats as the plain
    columns: class 0 = the events branch (seconds, t2 appended), class 1 = gala's own grid (Myr), then one class per
    distinct span shorter than dt (kicks.py:97)."""
    from ._lib import GRID_CLASS_DTYPE, pinned_empty
    n = t1_gyr.shape[0]
    t2_gyr, dt_myr = float(t2_gyr), float(dt_myr)
    span_gyr = t2_gyr - t1_gyr
    clip = np.flatnonzero(span_gyr < dt_myr * K.GYR_PER_MYR)
    rows = [(t2_gyr * K.SEC_PER_GYR, dt_myr * K.SEC_PER_MYR, K.MYR_PER_SEC, GRID_APPEND_T2),
            (t2_gyr * K.MYR_PER_GYR, dt_myr, 1.0, 0)]
    cls = pinned_empty(n, np.uint8)
    np.logical_not(kicked, out=cls.view(np.bool_))           # kicked -> 0, un-kicked -> 1
    if len(clip):
        key = np.stack([span_gyr[clip], kicked[clip].astype(np.float64)], axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        if 2 + len(uniq) > 256:
            return None
        for sp, kf in uniq:
            rows.append((t2_gyr * K.SEC_PER_GYR, sp * K.SEC_PER_GYR, K.MYR_PER_SEC, GRID_APPEND_T2) if kf else
                        (t2_gyr * K.MYR_PER_GYR, sp * K.MYR_PER_GYR, 1.0, 0))
        cls[clip] = 2 + np.asarray(inv).reshape(-1)
    classes = np.zeros(len(rows), dtype=GRID_CLASS_DTYPE)
    for i, (a, b, c, f) in enumerate(rows):
        classes[i] = (a, b, c, f, 0)
    return cls, classes


def build_orbit_batch(pos_kpc, vel_kms, tau_gyr, max_ev_time_gyr, timestep_myr, disrupted,
                      primary_events: Optional[EventTable] = None,
                      secondary_events: Optional[EventTable] = None, compact: bool = True,
                      threaded_rotation: bool = True) -> OrbitBatch:
    """Vectorised ``_iter_orbit_args``: primaries (population order) then disrupted secondaries.

    pos_kpc, vel_kms : (3, n) arrays; tau_gyr : (n,) look-back birth times; ``t1 = max_ev_time - tau``
    (pop.py:1180).  ``secondary_events.owner`` indexes the population (not the disrupted subset).

    compact : build the compact form of :class:`OrbitBatch` (what ``cw_integrate`` is sent), with the large arrays in
    page-locked memory; the plain arrays are then expanded only if something reads them.
    threaded_rotation : rotate the kicks of large batches on the library's host threads instead of numpy.
    """
    from ._lib import pinned_empty
    # (3, n) arrays, or three rows each (no stacking copy of a 10^7-system galaxy)
    pos_kpc = [np.asarray(r, dtype=np.float64) for r in pos_kpc]
    vel_kms = [np.asarray(r, dtype=np.float64) for r in vel_kms]
    tau_gyr = np.asarray(tau_gyr, dtype=np.float64)
    n = tau_gyr.shape[0]
    disrupted = np.asarray(disrupted, dtype=bool)
    dis_idx = np.flatnonzero(disrupted)
    n_dis = len(dis_idx)
    pe = primary_events if primary_events is not None else EventTable.empty()
    se = secondary_events if secondary_events is not None else EventTable.empty()
    n_orb = n + n_dis

    # events: primary rows keep their owner; secondary rows move to orbit n + rank(owner in disrupted)
    pe_off = pe.csr(n)
    rank = np.full(n, -1, dtype=np.int64)
    rank[dis_idx] = np.arange(n_dis)
    if len(se) and np.any(rank[se.owner] < 0):
        raise ValueError("secondary events given for a binary that is not disrupted")
    se_owner = rank[se.owner] if len(se) else se.owner
    se_off = EventTable(se_owner, se.tphys, se.dv, se.phase, se.inc).csr(n_dis)
    ev_off = np.concatenate([pe_off, pe_off[-1] + se_off[1:]])
    Kn = len(pe) + len(se)

    counts = ev_off[1:] - ev_off[:-1]
    # pop.py:1177-1183: `events` is None unless the binary has a supernova row; disrupted secondaries
    # always take the events branch (pop.py:1186-1193)
    kicked = counts > 0
    kicked[n:] = True
    t1_gyr_src = max_ev_time_gyr - tau_gyr

    ev_time = ev_dv = None
    if Kn:
        # (t1 + tphys * u.Myr) is formed in t1's unit (Gyr) and compared in seconds (kicks.py:134).  The primary and
        # the secondary rows are written into their own slices of the outputs: no concatenated copies of the tables.
        ev_time = pinned_empty(Kn)
        ev_dv = pinned_empty((Kn, 3))
        a = 0
        for tab in (pe, se):
            k = len(tab)
            if k == 0:
                continue
            et, ed = ev_time[a:a + k], ev_dv[a:a + k]
            np.multiply(np.asarray(tab.tphys, dtype=np.float64), K.GYR_PER_MYR, out=et)
            np.add(et, t1_gyr_src[np.asarray(tab.owner, dtype=np.int64)], out=et)     # population index of each event
            np.multiply(et, K.SEC_PER_GYR, out=et)
            dv = np.asarray(tab.dv, dtype=np.float64).reshape(-1, 3)
            ph, inc = np.asarray(tab.phase, dtype=np.float64), np.asarray(tab.inc, dtype=np.float64)
            if not (threaded_rotation and Kn >= 4096 and rotate_kicks_threaded(dv, ph, inc, K.KPC_MYR_PER_KMS, ed)):
                rotate_kicks(dv, ph, inc, out=ed)
                np.multiply(ed, K.KPC_MYR_PER_KMS, out=ed)
            a += k

    # the compact form needs every disrupted system's own orbit on the events branch (its start time is then the
    # same float in seconds for the system and its secondary) and few enough distinct grids
    grid = None
    if compact and n_orb < 2 ** 31 and Kn < 2 ** 31 and (n_dis == 0 or bool(kicked[dis_idx].all())):
        t1_gyr = np.concatenate([t1_gyr_src, t1_gyr_src[dis_idx]]) if n_dis else t1_gyr_src
        grid = _grid_classes(t1_gyr, max_ev_time_gyr, timestep_myr, kicked)
    if grid is not None:
        w0 = pinned_empty((6, n))
        for c in range(3):
            w0[c] = pos_kpc[c]
            np.multiply(vel_kms[c], K.KPC_MYR_PER_KMS, out=w0[3 + c])     # gala: km/s -> kpc/Myr on entry
        t1 = pinned_empty(n)
        np.multiply(t1_gyr_src, np.where(kicked[:n], K.SEC_PER_GYR, K.MYR_PER_GYR), out=t1)
        idx = pinned_empty(n_dis, np.int32)
        idx[:] = dis_idx
        off32 = None
        if Kn:
            off32 = pinned_empty(n_orb + 1, np.int32)
            off32[:] = ev_off
        c = dict(w0=w0, t1=t1, n_w0=n, w0_index=idx, grid_class=grid[0], classes=grid[1], ev_off32=off32)
        return OrbitBatch(ev_time=ev_time, ev_dv=ev_dv, n_primary=n, secondary_of=dis_idx, compact=c)

    src = np.concatenate([np.arange(n), dis_idx])          # population index of every orbit
    w0 = np.empty((6, n_orb))
    for c in range(3):
        w0[c] = pos_kpc[c][src]
        w0[3 + c] = vel_kms[c][src] * K.KPC_MYR_PER_KMS
    t1, t2, dt, to_myr, flags = _grid_columns(t1_gyr_src[src], max_ev_time_gyr, timestep_myr, kicked)
    return OrbitBatch(w0=w0, t1=t1, t2=t2, dt=dt, to_myr=to_myr, flags=flags,
                      ev_off=ev_off if Kn else None, ev_time=ev_time, ev_dv=ev_dv,
                      n_primary=n, secondary_of=dis_idx)
