"""The orbit-dependent helper of ``cogsworth.classify`` (reference ``src/cogsworth/classify.py:117-155``):
speeds of stars at the end of their orbits relative to the local circular velocity, used to label walkaways / runaways
(``classify.py:104-113``).  The reference loops over gala ``Orbit`` objects and calls ``potential.circular_velocity``;
here the final states come straight from the :class:`~cogsworth_b200.orbits.OrbitArray` and the circular velocity from
the ``cw_circular_velocity`` kernel -- the same parameter tables and device functions as the integrators.
"""
from __future__ import annotations

import numpy as np

from . import constants as K
from .orbits import OrbitArray

__all__ = ["get_rel_speed"]


def get_rel_speed(orbits, potential, context=None):
    """Drop-in for ``cogsworth.classify._get_rel_speed(orbits, potential)``: relative speed [km/s] of every orbit's last
    sample with respect to the circular velocity at that position.

    ``orbits`` : an :class:`OrbitArray` (or any sequence of orbit objects with ``[-1].pos.xyz`` / ``[-1].vel.d_xyz``).
    The expression is the reference's own, term for term (classify.py:147-153): cylindrical velocity components
    (v_R, v_T, v_z) and ``sqrt((v_R - v_circ)^2 + v_T^2 + v_z^2)`` -- the circular velocity is subtracted from the
    RADIAL component there, and a drop-in reproduces that."""
    from .kicks import _context
    if isinstance(orbits, OrbitArray):
        fp, fv = orbits.final_coords()
    else:
        fp = np.array([np.asarray(getattr(o[-1].pos.xyz, "value", o[-1].pos.xyz), dtype=np.float64).reshape(3) for o in orbits])
        fv = np.array([np.asarray(getattr(o[-1].vel.d_xyz, "value", o[-1].vel.d_xyz), dtype=np.float64).reshape(3)
                       for o in orbits]) * K.KMS_PER_KPC_MYR
    fp, fv = np.asarray(fp, dtype=np.float64).reshape(-1, 3), np.asarray(fv, dtype=np.float64).reshape(-1, 3)
    if len(fp) == 0:
        return np.zeros(0)
    ctx = _context(potential, context)
    v_circ = ctx.circular_velocity(np.ascontiguousarray(fp.T)) * K.KMS_PER_KPC_MYR
    x, y = fp[:, 0], fp[:, 1]
    rho = np.hypot(x, y)
    v_R = (x * fv[:, 0] + y * fv[:, 1]) / rho
    v_T = (x * fv[:, 1] - y * fv[:, 0]) / rho          # d_phi * rho
    v_z = fv[:, 2]
    return np.sqrt((v_R - v_circ) ** 2 + v_T ** 2 + v_z ** 2)
