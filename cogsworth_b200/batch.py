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

from dataclasses import dataclass, field
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


def rotate_kicks(dv_kms: np.ndarray, phase: np.ndarray, inc: np.ndarray) -> np.ndarray:
    """Vectorised ``get_kick_differential`` (kicks.py:42-46): BSE (vx,vy,vz) -> Galactocentric, km/s."""
    vx, vy, vz = dv_kms[:, 0], dv_kms[:, 1], dv_kms[:, 2]
    theta, phi = phase, inc
    v_X = vx * np.cos(theta) - vy * np.sin(theta) * np.cos(phi) + vz * np.sin(theta) * np.sin(phi)
    v_Y = vx * np.sin(theta) + vy * np.cos(theta) * np.cos(phi) - vz * np.cos(theta) * np.sin(phi)
    v_Z = vy * np.sin(phi) + vz * np.cos(phi)
    return np.stack([v_X, v_Y, v_Z], axis=1)


@dataclass
class OrbitBatch:
    """Everything ``cw_integrate`` needs for a set of orbits (see include/cogsworth_b200.h)."""
    w0: np.ndarray           # (6, n)  kpc, kpc/Myr
    t1: np.ndarray           # grid unit
    t2: np.ndarray
    dt: np.ndarray
    to_myr: np.ndarray
    flags: np.ndarray        # int32
    ev_off: Optional[np.ndarray] = None   # (n+1,) int64
    ev_time: Optional[np.ndarray] = None  # (K,) grid unit
    ev_dv: Optional[np.ndarray] = None    # (K, 3) kpc/Myr
    n_primary: int = 0       # the first n_primary orbits are bound systems / primaries (Q7)
    secondary_of: Optional[np.ndarray] = None  # population index of each disrupted secondary
    meta: dict = field(default_factory=dict)

    @property
    def n_orbits(self) -> int:
        return self.w0.shape[1]

    @property
    def n_events(self) -> int:
        return 0 if self.ev_off is None else int(self.ev_off[-1])

    def take(self, idx) -> "OrbitBatch":
        """Sub-batch (used by the retry loop for failed orbits)."""
        idx = np.asarray(idx, dtype=np.int64)
        ev_off = ev_time = ev_dv = None
        if self.ev_off is not None:
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


def build_orbit_batch(pos_kpc, vel_kms, tau_gyr, max_ev_time_gyr, timestep_myr, disrupted,
                      primary_events: Optional[EventTable] = None,
                      secondary_events: Optional[EventTable] = None) -> OrbitBatch:
    """Vectorised ``_iter_orbit_args``: primaries (population order) then disrupted secondaries.

    pos_kpc, vel_kms : (3, n) arrays; tau_gyr : (n,) look-back birth times; ``t1 = max_ev_time - tau``
    (pop.py:1180).  ``secondary_events.owner`` indexes the population (not the disrupted subset).
    """
    pos_kpc = np.asarray(pos_kpc, dtype=np.float64)
    vel_kms = np.asarray(vel_kms, dtype=np.float64)
    tau_gyr = np.asarray(tau_gyr, dtype=np.float64)
    n = tau_gyr.shape[0]
    disrupted = np.asarray(disrupted, dtype=bool)
    dis_idx = np.flatnonzero(disrupted)
    n_dis = len(dis_idx)
    pe = primary_events if primary_events is not None else EventTable.empty()
    se = secondary_events if secondary_events is not None else EventTable.empty()

    src = np.concatenate([np.arange(n), dis_idx])          # population index of every orbit
    n_orb = n + n_dis
    w0 = np.empty((6, n_orb))
    w0[:3] = pos_kpc[:, src]
    w0[3:] = vel_kms[:, src] * K.KPC_MYR_PER_KMS            # gala: km/s -> kpc/Myr on entry

    # events: primary rows keep their owner; secondary rows move to orbit n + rank(owner in disrupted)
    pe_off = pe.csr(n)
    rank = np.full(n, -1, dtype=np.int64)
    rank[dis_idx] = np.arange(n_dis)
    if len(se) and np.any(rank[se.owner] < 0):
        raise ValueError("secondary events given for a binary that is not disrupted")
    se_owner = rank[se.owner] if len(se) else se.owner
    se_tab = EventTable(se_owner, se.tphys, se.dv, se.phase, se.inc)
    se_off = se_tab.csr(n_dis)
    ev_off = np.concatenate([pe_off, pe_off[-1] + se_off[1:]])
    tphys = np.concatenate([pe.tphys, se.tphys]).astype(np.float64)
    dv = np.concatenate([pe.dv.reshape(-1, 3), se.dv.reshape(-1, 3)]).astype(np.float64)
    phase = np.concatenate([pe.phase, se.phase]).astype(np.float64)
    inc = np.concatenate([pe.inc, se.inc]).astype(np.float64)
    owner_orbit = np.concatenate([pe.owner, n + se_owner]).astype(np.int64)

    counts = ev_off[1:] - ev_off[:-1]
    # pop.py:1177-1183: `events` is None unless the binary has a supernova row; disrupted secondaries
    # always take the events branch (pop.py:1186-1193)
    kicked = counts > 0
    kicked[n:] = True
    t1_gyr = max_ev_time_gyr - tau_gyr[src]
    t1, t2, dt, to_myr, flags = _grid_columns(t1_gyr, max_ev_time_gyr, timestep_myr, kicked)

    Kn = len(tphys)
    ev_time = ev_dv = None
    if Kn:
        # (t1 + tphys * u.Myr) is formed in t1's unit (Gyr) and compared in seconds (kicks.py:134)
        ev_time = (t1_gyr[owner_orbit] + tphys * K.GYR_PER_MYR) * K.SEC_PER_GYR
        ev_dv = rotate_kicks(dv, phase, inc) * K.KPC_MYR_PER_KMS
    return OrbitBatch(w0=w0, t1=t1, t2=t2, dt=dt, to_myr=to_myr, flags=flags,
                      ev_off=ev_off if Kn else None, ev_time=ev_time, ev_dv=ev_dv,
                      n_primary=n, secondary_of=dis_idx)
