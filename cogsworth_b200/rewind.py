"""Backward integration for ``cogsworth.hydro.rewind`` (reference ``src/cogsworth/hydro/rewind.py``).

``rewind_to_formation`` fans ``_tohspans(pot, wf, t1, t2, dt)`` -- gala ``integrate_orbit(wf, t1=final_time,
t2=t_form, dt=-1 Myr, Integrator=DOPRI853, store_all=False)[-1]`` -- out over a multiprocessing pool, one star
particle per task (``rewind.py:16-19,57-63``).  :func:`rewind_orbits` is the batch form of that fan-out on
the GPU: one ``cw_integrate`` call with backward grids (``t2 < t1``, ``dt < 0``; gala's own grid in Myr, ``t2``
appended as the last node -- the backward branch of ``parse_time_specification`` [gala-recall]).

Only analytic potentials (``cogsworth_b200.potential``) have kernels; the SCF / agama fits that
``HydroPopulation`` builds from a snapshot are out of scope and raise ``NotImplementedError`` in
``potential.from_gala``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import constants as K
from .kicks import _context, resolve_integrator

__all__ = ["rewind_orbits"]


def rewind_orbits(pos_kpc, vel_kms, t1_myr, t2_myr, dt_myr=-1.0, potential=None, integrator=None,
                  integrator_kwargs=None, context=None):
    """Integrate ``n`` particles from ``t1_myr`` back to their own ``t2_myr[i]`` (``dt_myr < 0``).

    pos_kpc, vel_kms : (3, n) present-day Galactocentric positions [kpc] / velocities [km/s]
    t1_myr : scalar or (n,) start time (the snapshot time);  t2_myr : (n,) formation times (< t1)
    Returns (pos (n, 3) kpc, vel (n, 3) km/s, ok (n,) bool) -- the ``w0[i].xyz / .v_xyz`` the reference
    collects at ``rewind.py:79-81``; rows of failed integrations are ``inf``.
    """
    from .potential import MilkyWayPotential
    pos = np.ascontiguousarray(pos_kpc, dtype=np.float64).reshape(3, -1)
    vel = np.ascontiguousarray(vel_kms, dtype=np.float64).reshape(3, -1)
    n = pos.shape[1]
    if not dt_myr < 0:
        raise ValueError("rewind needs dt < 0 (gala: 'If t2 < t1, dt must be negative')")
    t1 = np.broadcast_to(np.asarray(t1_myr, dtype=np.float64), (n,)).copy()
    t2 = np.broadcast_to(np.asarray(t2_myr, dtype=np.float64), (n,)).copy()
    if np.any(t2 >= t1):
        raise ValueError("every formation time must lie before the snapshot time")
    dt = np.maximum(np.full(n, float(dt_myr)), t2 - t1)         # |dt| no larger than the span
    w0 = np.vstack([pos, vel * K.KPC_MYR_PER_KMS])
    ctx = _context(MilkyWayPotential("v2") if potential is None else potential, context)
    kw = dict(integrator_kwargs or {})
    res = ctx.integrate(w0, t1, t2, dt, np.ones(n), np.full(n, _lib.GRID_APPEND_T2, dtype=np.int32),
                        integrator=resolve_integrator(integrator, kw), atol=kw.get("atol", 1e-10),
                        rtol=kw.get("rtol", 1e-10), nmax=kw.get("nmax", 0) or 0)
    ok = res["status"] == 0
    wf = res["w_final"]
    out_pos = np.where(ok[:, None], wf[:3].T, np.inf)
    out_vel = np.where(ok[:, None], wf[3:].T * K.KMS_PER_KPC_MYR, np.inf)
    return out_pos, out_vel, ok
