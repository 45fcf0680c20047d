"""Orbit results without one Python object per system.

The reference returns a numpy object array of gala ``Orbit`` s (``src/cogsworth/pop.py:1313``) and
loops over it in Python to get ``final_pos`` / ``final_vel`` (``pop.py:1338-1348``); at 10^6-10^7
orbits that alone costs minutes.  :class:`OrbitArray` keeps the kernels' structure-of-arrays output
-- the very layout ``Population.save`` writes to HDF5 (``orbits/{offsets,pos,vel,t}``,
``pop.py:1853-1876``) -- and materialises an orbit object only when one is indexed.  The
materialised object is a real ``gala.dynamics.Orbit`` when gala is importable, otherwise the light
:class:`Orbit` below, which exposes the attribute surface the reference uses downstream
(``.pos.xyz``, ``.vel.d_xyz``, ``.t``, ``.x/.y/.z``, ``[-1]``, ``[-1:]``, masks, ``.shape``).
"""
from __future__ import annotations

import numpy as np

from . import constants as K

__all__ = ["Orbit", "OrbitArray"]


def _have_gala():
    try:
        import gala.dynamics  # noqa: F401
        import astropy.units  # noqa: F401
        return True
    except Exception:
        return False


class _Vec3:
    """Tiny stand-in for an astropy Cartesian representation / differential."""

    def __init__(self, a, prefix):
        self._a = a
        self._p = prefix

    @property
    def xyz(self):
        return self._a

    d_xyz = xyz

    def __getattr__(self, name):
        idx = {"x": 0, "y": 1, "z": 2, "d_x": 0, "d_y": 1, "d_z": 2}.get(name)
        if idx is None:
            raise AttributeError(name)
        return self._a[idx]

    def __len__(self):
        return self._a.shape[1]


class Orbit:
    """pos (3, L) kpc, vel (3, L) kpc/Myr, t (L,) Myr -- gala ``Orbit`` look-alike."""

    def __init__(self, pos, vel, t):
        self._pos = np.atleast_2d(np.asarray(pos, dtype=np.float64))
        self._vel = np.atleast_2d(np.asarray(vel, dtype=np.float64))
        self.t = np.atleast_1d(np.asarray(t, dtype=np.float64))

    @property
    def pos(self):
        return _Vec3(self._pos, "")

    @property
    def vel(self):
        return _Vec3(self._vel, "d_")

    x = property(lambda s: s._pos[0])
    y = property(lambda s: s._pos[1])
    z = property(lambda s: s._pos[2])
    v_x = property(lambda s: s._vel[0])
    v_y = property(lambda s: s._vel[1])
    v_z = property(lambda s: s._vel[2])

    @property
    def shape(self):
        return (self._pos.shape[1],)

    @property
    def ndim(self):
        return 3

    def __len__(self):
        return self._pos.shape[1]

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            key = slice(key, key + 1 if key != -1 else None)
        return Orbit(self._pos[:, key], self._vel[:, key], self.t[key])

    def w(self):
        return np.vstack([self._pos, self._vel])

    # the slice of gala's representation API the reference touches (classify.py:137-153, plot.py:541-577)
    class _Cyl:
        def __init__(self, pos):
            self.rho = np.hypot(pos[0], pos[1])
            self.phi = np.arctan2(pos[1], pos[0])
            self.z = pos[2]

    @property
    def cylindrical(self):
        return Orbit._Cyl(self._pos)

    def represent_as(self, name):
        if str(name).lower().startswith("cyl"):
            return self.cylindrical
        if str(name).lower().startswith("cart"):
            return self
        raise NotImplementedError(f"representation {name!r} needs gala (install it to get real gala Orbit objects)")

    @property
    def data(self):
        return self.pos

    def __repr__(self):
        return f"<Orbit cartesian, dim=3, shape={self.shape}>"


class OrbitArray:
    """Lazy, ndarray-like container of orbits backed by SoA trajectory buffers.

    offsets (n+1,), pos (3, total) kpc, vel (3, total) kpc/Myr, t (total,) Myr; ``ok[i]`` False marks
    a failed integration (the reference stores ``None`` there).  Supports ``len``, integer / slice /
    boolean / fancy indexing and ``np.concatenate``-style joining via :meth:`concatenate`.
    """

    def __init__(self, offsets, pos, vel, t, ok=None, index=None):
        # offsets = None: one sample per orbit (final-state runs) -- nothing of size n is built until it is asked for
        self._offsets = None if offsets is None else np.asarray(offsets, dtype=np.int64)
        self.pos, self.vel = pos, vel
        self._t = t                                  # array, or a callable that makes it (labels are rarely read)
        self._n = pos.shape[1] if offsets is None else len(self._offsets) - 1
        self._ok = None if ok is None else np.asarray(ok, dtype=bool)     # None = every orbit succeeded
        self._idx = None if index is None else np.asarray(index, dtype=np.int64)

    @classmethod
    def from_final(cls, w_final, t_final, ok=None):
        """Final-state-only result (store_all=False): every orbit has exactly one sample.  ``t_final`` may be a
        callable returning the (n,) array of time labels."""
        return cls(None, w_final[:3], w_final[3:], t_final if callable(t_final) else np.asarray(t_final), ok)

    @property
    def offsets(self):
        if self._offsets is None:
            self._offsets = np.arange(self._n + 1, dtype=np.int64)
        return self._offsets

    @property
    def t(self):
        if callable(self._t):
            self._t = np.asarray(self._t())
        return self._t

    @property
    def ok(self):
        if self._ok is None:
            self._ok = np.ones(self._n, dtype=bool)
        return self._ok

    @property
    def _index(self):
        if self._idx is None:
            self._idx = np.arange(self._n)
        return self._idx

    @property
    def n_failed(self) -> int:
        """Number of selected orbits whose integration failed (``None`` entries of the reference's array)."""
        if self._ok is None:
            return 0
        return int((~self._ok[self._index]).sum()) if self._idx is not None else int((~self._ok).sum())

    # -- container protocol --------------------------------------------------------------------
    def __len__(self):
        return self._n if self._idx is None else len(self._idx)

    @property
    def shape(self):
        return (len(self),)

    def _materialise(self, i):
        if self._ok is not None and not self._ok[i]:
            return None
        a, b = (i, i + 1) if self._offsets is None else (self.offsets[i], self.offsets[i + 1])
        pos, vel, t = self.pos[:, a:b], self.vel[:, a:b], self.t[a:b]
        if _have_gala():
            import astropy.units as u
            import gala.dynamics as gd
            return gd.Orbit(pos * u.kpc, vel * u.kpc / u.Myr, t=t * u.Myr)
        return Orbit(pos, vel, t)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self._materialise(int(key) % self._n if self._idx is None else self._index[key])
        sub = self._index[key]
        return OrbitArray(self._offsets, self.pos, self.vel, self._t, self._ok, np.atleast_1d(sub))

    def __iter__(self):
        for i in self._index:
            yield self._materialise(i)

    def to_object_array(self):
        """The reference's representation: ``np.array(orbits, dtype=object)`` (pop.py:1313)."""
        out = np.empty(len(self), dtype=object)
        for k, i in enumerate(self._index):
            out[k] = self._materialise(i)
        return out

    def __array__(self, dtype=None, copy=None):
        return self.to_object_array()

    # -- bulk accessors (no per-orbit objects) ---------------------------------------------------
    def final_coords(self):
        """(final_pos [kpc], final_vel [km/s]) of shape (n, 3); ``inf`` rows for failed orbits
        (pop.py:1338-1348)."""
        if self._offsets is None and self._idx is None:       # one sample per orbit: no gather at all
            fp, fv = self.pos.T, self.vel.T * K.KMS_PER_KPC_MYR
            if self._ok is not None and not self._ok.all():
                fp, fv = np.where(self._ok[:, None], fp, np.inf), np.where(self._ok[:, None], fv, np.inf)
            return fp, fv
        last = self.offsets[self._index + 1] - 1
        ok = self.ok[self._index] & (last >= self.offsets[self._index])
        last = np.where(ok, last, 0)
        fp = np.where(ok[:, None], self.pos[:, last].T, np.inf)
        fv = np.where(ok[:, None], self.vel[:, last].T * K.KMS_PER_KPC_MYR, np.inf)
        return fp, fv

    def lengths(self):
        return (self.offsets[1:] - self.offsets[:-1])[self._index]

    def compact(self):
        """(offsets, pos, vel_kms, t) of the selected orbits, contiguous -- the arrays
        ``Population.save`` writes under ``orbits/`` (pop.py:1853-1876)."""
        lens = self.lengths()
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        cols = np.concatenate([np.arange(self.offsets[i], self.offsets[i + 1]) for i in self._index]) \
            if len(lens) else np.zeros(0, dtype=np.int64)
        return off, self.pos[:, cols], self.vel[:, cols] * K.KMS_PER_KPC_MYR, self.t[cols]

    @staticmethod
    def concatenate(arrays):
        offs, poss, vels, ts, oks = [np.zeros(1, dtype=np.int64)], [], [], [], []
        base = 0
        for a in arrays:
            off, pos, vel_kms, t = a.compact()
            offs.append(off[1:] + base)
            base += off[-1]
            poss.append(pos); vels.append(vel_kms / K.KMS_PER_KPC_MYR); ts.append(t)
            oks.append(a.ok[a._index])
        return OrbitArray(np.concatenate(offs), np.concatenate(poss, axis=1), np.concatenate(vels, axis=1),
                          np.concatenate(ts), np.concatenate(oks))
