"""Drop-in for ``cogsworth.kicks`` (reference ``src/cogsworth/kicks.py``) on the B200 path.

``integrate_orbit_with_events`` keeps the reference signature and semantics (time grid, kick on
the last grid node strictly before the event, segment restart, retry with a smaller time-step) but
the arithmetic runs in the CUDA kernels behind ``cw_integrate``; nothing here integrates on the
CPU.  ``integrate_orbits`` is the batch form the population driver uses.

Without astropy (it is not a dependency of this package) quantities are plain numbers in the
units ``Population`` passes: positions kpc, velocities km/s, ``t1`` / ``t2`` Gyr, ``dt`` Myr,
``tphys`` Myr, kicks km/s.  astropy Quantities are accepted and converted when astropy is present.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import constants as K
from .batch import EventTable, OrbitBatch, build_orbit_batch, rotate_kicks
from .orbits import Orbit, OrbitArray
from .potential import CompositePotential, MilkyWayPotential, from_gala

__all__ = ["get_kick_differential", "integrate_orbit_with_events", "integrate_orbits",
           "resolve_integrator", "IntegrationFailed"]

# statuses that correspond to gala's RuntimeError("Integration failed ...") -> retry / None
_FAILED = (-2, -3, -4, -5)


class IntegrationFailed(RuntimeError):
    pass


def _value(q, unit):
    """Plain float(s) of ``q`` in ``unit`` (an astropy unit name); numbers pass through."""
    if hasattr(q, "unit"):
        import astropy.units as u
        return q.to(getattr(u, unit) if "/" not in unit else u.Unit(unit)).value
    return q


def get_kick_differential(delta_v_sys_x, delta_v_sys_y, delta_v_sys_z, phase=None, inclination=None):
    """Rotate a BSE-frame systemic kick into the Galactocentric frame (kicks.py:12-49).

    Returns (v_X, v_Y, v_Z) in the input units -- an astropy ``CartesianDifferential`` if the inputs
    are Quantities, else a length-3 ndarray.  Angles are drawn from the global numpy RNG when not
    given, as in the reference."""
    theta = np.random.uniform(0, 2 * np.pi) if phase is None else phase
    phi = np.arccos(2 * np.random.rand() - 1.0) if inclination is None else inclination
    is_q = hasattr(delta_v_sys_x, "unit")
    vx, vy, vz = delta_v_sys_x, delta_v_sys_y, delta_v_sys_z
    v_X = vx * np.cos(theta) - vy * np.sin(theta) * np.cos(phi) + vz * np.sin(theta) * np.sin(phi)
    v_Y = vx * np.sin(theta) + vy * np.cos(theta) * np.cos(phi) - vz * np.cos(theta) * np.sin(phi)
    v_Z = vy * np.sin(phi) + vz * np.cos(phi)
    if is_q:
        import astropy.coordinates as coords
        return coords.CartesianDifferential(v_X, v_Y, v_Z)
    return np.array([v_X, v_Y, v_Z], dtype=np.float64)


#: How ``gi.DOPRI853Integrator`` is executed when the caller does not say: "restart" = a fresh
#: dop853() on every grid interval (gala's classic dop853_integrate_hamiltonian), "dense" = one call
#: per segment with Hairer's continuous extension at the grid nodes.  Per call:
#: ``integrator="dop853_dense"`` or ``integrator_kwargs={"dense_output": True}``.
DOP853_MODE = "restart"


def resolve_integrator(integrator, integrator_kwargs=None):
    """Map a gala Integrator class / instance / name onto the kernel id (pop.py:130, kicks.py:54)."""
    dense = (integrator_kwargs or {}).get("dense_output")
    if dense is None:
        dense = DOP853_MODE == "dense"
    dop = _lib.DOP853_DENSE if dense else _lib.DOP853
    if integrator is None:
        return dop
    if isinstance(integrator, (int, np.integer)):
        return int(integrator)
    name = integrator if isinstance(integrator, str) else getattr(integrator, "__name__", type(integrator).__name__)
    name = name.lower()
    if "leapfrog" in name:
        return _lib.LEAPFROG
    if "dop" in name or "853" in name:
        if "dense" in name:
            return _lib.DOP853_DENSE
        if "restart" in name:
            return _lib.DOP853
        return dop
    raise NotImplementedError(f"integrator {integrator!r} has no sm_100a kernel (Leapfrog, DOPRI853 only)")


def _context(potential, context=None):
    ctx = context if context is not None else _lib.default_context()
    pot = from_gala(potential) if not isinstance(potential, CompositePotential) else potential
    # compared by value: from_gala returns a fresh object on every call
    if ctx.potential is not pot and (ctx.potential is None or ctx.potential.fingerprint() != pot.fingerprint()):
        ctx.set_potential(pot)
    return ctx


def _grid_times_myr(batch: OrbitBatch, n_nodes):
    """Orbit.t of final-state-only results: the time of the last node (Myr).  For appended grids that
    is t2; otherwise t1 + (n-1) dt ACCUMULATED as the grid is -- only the label, the kernels never read it."""
    app = (batch.flags & 1) == 1
    t_last = np.where(app, batch.t2, batch.t1)
    idx = np.flatnonzero(~app)
    if len(idx):
        # one vector addition per step over the orbits that still have steps left (they are sorted by length first)
        steps = np.maximum(np.asarray(n_nodes)[idx].astype(np.int64) - 1, 0)
        order = np.argsort(-steps, kind="stable")
        idx, steps = idx[order], steps[order]
        t, dt = batch.t1[idx].copy(), batch.dt[idx]
        for j in range(int(steps[0]) if len(steps) else 0):
            live = int(np.searchsorted(-steps, -j, side="left"))     # orbits with steps > j
            t[:live] += dt[:live]
        t_last[idx] = t
    return t_last * batch.to_myr


def _ranges(starts, lens):
    """Concatenation of arange(s, s + l) for every (s, l)."""
    lens = np.asarray(lens, dtype=np.int64)
    tot = int(lens.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.cumsum(lens)
    return np.repeat(np.asarray(starts, dtype=np.int64) - (ends - lens), lens) + np.arange(tot, dtype=np.int64)


def integrate_orbits(batch: OrbitBatch, potential=None, store_all=True, integrator=None, integrator_kwargs=None,
                     max_retries=2, timestep_multiplier=0.1, context=None, final_times=False):
    """Integrate every orbit of ``batch`` on the GPU(s); the batch form of
    ``integrate_orbit_with_events`` incl. its retry loop (kicks.py:113,193-203).

    Returns (OrbitArray, info) where ``info`` has ``status``, ``n_nodes``, ``ev_step``, ``stats``."""
    potential = MilkyWayPotential("v2") if potential is None else potential
    ctx = _context(potential, context)
    kw = dict(integrator_kwargs or {})
    integ = resolve_integrator(integrator, kw)
    common = dict(integrator=integ, atol=kw.get("atol", 1e-10), rtol=kw.get("rtol", 1e-10),
                  nmax=kw.get("nmax", 0) or 0)

    def run(b: OrbitBatch):
        return ctx.integrate_batch(b, store_all=store_all, **common)

    res = run(batch)
    status = res["status"]
    n = batch.n_orbits
    clean = res.get("n_failed", -1) == 0           # nothing failed: no per-orbit scan of a 10^7-orbit batch below
    if not clean and np.any(status == -1):
        bad = np.flatnonzero(status == -1)
        raise ValueError(f"{len(bad)} orbit(s) have inconsistent events/time grid (first: orbit {bad[0]}); "
                         "the reference fails to stitch such orbits (kicks.py:187-189)")

    # trajectories per orbit may be replaced by a retry with a finer grid -> keep a list of pieces
    pieces = None
    if store_all:
        pieces = dict(off=res["traj_off"], pos=res["traj_pos"], vel=res["traj_vel"], t=res["traj_t"])
    w_final, n_nodes = res["w_final"], res["n_nodes"]
    retry_results = {}
    # retries only exist on the events branch, max_retries attempts in total (reference quirk Q4)
    kicked = None
    dt_scale = 1.0
    for _ in range(0 if clean else max(int(max_retries) - 1, 0)):
        failed = np.flatnonzero(status != 0)
        if len(failed) == 0:
            break
        if kicked is None:
            kicked = (batch.flags & 1) == 1
        failed = failed[np.isin(status[failed], _FAILED) & kicked[failed]]
        if len(failed) == 0:
            break
        dt_scale *= timestep_multiplier
        sub = batch.take(failed)
        sub.dt = sub.dt * dt_scale
        r = run(sub)
        for k, i in enumerate(failed):
            status[i] = r["status"][k]
            if r["status"][k] == 0:
                w_final[:, i] = r["w_final"][:, k]
                n_nodes[i] = r["n_nodes"][k]
                if store_all:
                    a, b = r["traj_off"][k], r["traj_off"][k + 1]
                    retry_results[i] = (r["traj_pos"][:, a:b], r["traj_vel"][:, a:b], r["traj_t"][a:b])
    ok = None if clean else status == 0
    if store_all:
        off, pos, vel, t = pieces["off"], pieces["pos"], pieces["vel"], pieces["t"]
        if retry_results:   # rebuild the SoA buffers with the re-gridded orbits spliced in
            lens_old = off[1:] - off[:-1]
            lens = lens_old.copy()
            for i, (p_, v_, t_) in retry_results.items():
                lens[i] = p_.shape[1]
            noff = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=noff[1:])
            npos, nvel, nt = np.empty((3, noff[-1])), np.empty((3, noff[-1])), np.empty(noff[-1])
            keep = np.ones(n, dtype=bool)
            keep[list(retry_results)] = False
            src, dst = _ranges(off[:-1][keep], lens_old[keep]), _ranges(noff[:-1][keep], lens[keep])
            npos[:, dst], nvel[:, dst], nt[dst] = pos[:, src], vel[:, src], t[src]
            for i, (p_, v_, t_) in retry_results.items():
                npos[:, noff[i]:noff[i + 1]], nvel[:, noff[i]:noff[i + 1]], nt[noff[i]:noff[i + 1]] = p_, v_, t_
            off, pos, vel, t = noff, npos, nvel, nt
        orbits = OrbitArray(off, pos, vel, t, ok)
    else:
        if final_times:
            t_final = _grid_times_myr(batch, n_nodes)
        else:       # the label of the one stored sample: t2 on the events branch (appended), unknown without a walk
            def t_final():
                return np.where((batch.flags & 1) == 1, batch.t2 * batch.to_myr, np.nan)
        orbits = OrbitArray.from_final(w_final, t_final, ok)
    info = dict(status=status, n_nodes=n_nodes, ev_step=res["ev_step"], stats=res["stats"], w_final=w_final)
    return orbits, info


def integrate_orbit_with_events(w0, t1, t2, dt, potential=None, events=None, store_all=True,
                                integrator=None, integrator_kwargs={}, max_retries=2,
                                timestep_multiplier=0.1, context=None):
    """Drop-in for ``cogsworth.kicks.integrate_orbit_with_events`` (kicks.py:52-209), one orbit.

    w0 : gala ``PhaseSpacePosition`` or a length-6 sequence (x, y, z [kpc], v_x, v_y, v_z [km/s]).
    t1, t2 : start / end time (Quantity or number in Gyr); dt : time-step (Quantity or number in Myr).
    events : ``None`` or a DataFrame / list of dicts with ``tphys`` [Myr], ``delta_vsys_x|y|z`` [km/s],
    ``phase``, ``inc``.  Returns an orbit object (gala ``Orbit`` when gala is installed), or ``None``
    if the integration failed after the retries.
    """
    if hasattr(w0, "pos") and hasattr(w0, "vel"):
        pos = np.array([_value(getattr(w0.pos, c), "kpc") for c in "xyz"], dtype=np.float64).reshape(3)
        vel = np.array([_value(getattr(w0.vel, c), "km/s") for c in ("d_x", "d_y", "d_z")],
                       dtype=np.float64).reshape(3)
    else:
        w = np.asarray(w0, dtype=np.float64).reshape(6)
        pos, vel = w[:3], w[3:]
    t1_gyr, t2_gyr = float(_value(t1, "Gyr")), float(_value(t2, "Gyr"))
    dt_myr = float(_value(dt, "Myr"))

    prim = None
    if events is not None:
        if hasattr(events, "iterrows"):
            rows = [r for _, r in events.iterrows()]
        else:
            rows = list(events)
        k = len(rows)
        prim = EventTable(np.zeros(k, dtype=np.int64),
                          np.array([float(r["tphys"]) for r in rows]),
                          np.array([[float(r["delta_vsys_x"]), float(r["delta_vsys_y"]), float(r["delta_vsys_z"])]
                                    for r in rows]).reshape(k, 3),
                          np.array([float(r["phase"]) for r in rows]),
                          np.array([float(r["inc"]) for r in rows]))
    batch = build_orbit_batch(pos.reshape(3, 1), vel.reshape(3, 1), np.array([t2_gyr - t1_gyr]), t2_gyr,
                              dt_myr, np.array([False]), prim, None)
    # t1 must be the caller's value, not t2 - (t2 - t1)
    from .batch import _grid_columns
    kicked = np.array([events is not None])
    batch.t1, batch.t2, batch.dt, batch.to_myr, batch.flags = _grid_columns(
        np.array([t1_gyr]), t2_gyr, dt_myr, kicked)
    if events is not None and len(prim):
        batch.ev_time = (t1_gyr + prim.tphys * K.GYR_PER_MYR) * K.SEC_PER_GYR
    orbits, info = integrate_orbits(batch, potential=potential, store_all=store_all, integrator=integrator,
                                    integrator_kwargs=integrator_kwargs, max_retries=max_retries,
                                    timestep_multiplier=timestep_multiplier, context=context,
                                    final_times=True)
    return orbits[0]
