"""Orbit wire format of ``Population.save`` / ``Population.orbits`` (reference
``src/cogsworth/pop.py:1853-1876`` and ``:696-719``) written and read straight
from the structure-of-arrays trajectory buffers the kernels fill.

The reference builds the four arrays of the ``orbits`` group with a Python loop over gala ``Orbit``
objects (``pop.py:1866-1870``) and rebuilds one ``Orbit`` per system on load (``pop.py:711-713``).
Here the kernels' output already IS that layout::

    orbits/offsets  int64  (n + 1,)       cumulative sample counts
    orbits/pos      f64    (3, total)     kpc
    orbits/vel      f64    (3, total)     km/s      (kernel buffers are kpc/Myr: one multiply)
    orbits/t        f64    (total,)       Myr

so saving is O(bytes) and loading returns a lazy :class:`~cogsworth_b200.orbits.OrbitArray`.

HDF5 files are written / read with ``h5py`` when it is importable (cogsworth's own dependency) and otherwise with
:mod:`cogsworth_b200.h5mini`, a ``struct``-only implementation of exactly the subset of the format these four arrays
need (superblock 0, symbol-table groups, contiguous datasets -- h5py's defaults), so the reference's own container
works in this image too; ``.npz`` file names select a dependency-free numpy container with the keys
``orbits/offsets`` ... instead.  Without h5py an ``.h5`` file can only be created, not appended to.
"""
from __future__ import annotations

import os

import numpy as np

from . import constants as K
from .orbits import OrbitArray

__all__ = ["save_orbits", "load_orbits", "final_coords_from_file"]

_KEYS = ("offsets", "pos", "vel", "t")


def _is_hdf5(file_name: str) -> bool:
    return os.path.splitext(file_name)[1].lower() in (".h5", ".hdf5", ".hdf")


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:           # pragma: no cover - depends on the environment
        return None


def save_orbits(file_name: str, orbits: OrbitArray) -> None:
    """Write ``orbits`` under the ``orbits`` group of ``file_name`` (appending to an existing
    population file, as ``Population.save`` does at pop.py:1872-1876).  Failed orbits (``None`` in
    the reference) must have been dropped first, as ``perform_galactic_evolution`` does
    (pop.py:1269-1311): the reference format cannot represent them."""
    if not np.all(orbits.ok[orbits._index]):
        raise ValueError("cannot save failed orbits (None in the reference); drop them first")
    off, pos, vel_kms, t = orbits.compact()
    data = {"offsets": off, "pos": np.ascontiguousarray(pos), "vel": np.ascontiguousarray(vel_kms),
            "t": np.ascontiguousarray(t)}
    if _is_hdf5(file_name):
        h5 = _h5py()
        if h5 is None:
            from . import h5mini
            if os.path.exists(file_name):
                raise FileExistsError(f"{file_name} exists: appending the orbits group to a population file needs h5py "
                                      "(cogsworth_b200.h5mini only creates files)")
            h5mini.write_datasets(file_name, "orbits", {k: data[k] for k in _KEYS})
            return
        with h5.File(file_name, "a") as f:
            if "orbits" in f:
                del f["orbits"]
            g = f.create_group("orbits")
            for k in _KEYS:
                g[k] = data[k]
        return
    if not file_name.endswith(".npz"):
        raise ValueError("file name must end in .h5 / .hdf5 (HDF5, needs h5py) or .npz")
    existing = {}
    if os.path.exists(file_name):
        with np.load(file_name) as z:
            existing = {k: z[k] for k in z.files if not k.startswith("orbits/")}
    existing.update({f"orbits/{k}": v for k, v in data.items()})
    np.savez(file_name, **existing)


def _read(file_name: str):
    if _is_hdf5(file_name):
        h5 = _h5py()
        if h5 is None:
            from . import h5mini
            try:
                d = h5mini.read_datasets(file_name, "orbits", _KEYS)
            except KeyError:
                raise ValueError(f"No orbits found in population file ({file_name})") from None   # pop.py:704-705
            return tuple(d[k] for k in _KEYS)
        with h5.File(file_name, "r") as f:
            if "orbits" not in f:
                raise ValueError(f"No orbits found in population file ({file_name})")   # pop.py:704-705
            return tuple(f["orbits"][k][...] for k in _KEYS)
    with np.load(file_name) as z:
        if "orbits/offsets" not in z.files:
            raise ValueError(f"No orbits found in population file ({file_name})")
        return tuple(z[f"orbits/{k}"] for k in _KEYS)


def load_orbits(file_name: str) -> OrbitArray:
    """Lazy orbits of a saved population (pop.py:696-719) -- no per-system objects are built."""
    off, pos, vel_kms, t = _read(file_name)
    return OrbitArray(off, pos, vel_kms / K.KMS_PER_KPC_MYR, t)


def final_coords_from_file(file_name: str):
    """(final_pos [kpc], final_vel [km/s]), each (n, 3): what ``Population.orbits`` derives while
    loading (pop.py:715-718) -- here without touching the trajectories beyond their last columns."""
    off, pos, vel_kms, _ = _read(file_name)
    last = off[1:] - 1
    return pos[:, last].T, vel_kms[:, last].T
