"""``Population.perform_galactic_evolution`` on the B200 path.

The reference swaps stages by subclassing ``Population`` and overriding the stage method
(``COMPASPopulation.perform_stellar_evolution`` ``src/cogsworth/interop/compas/pop.py:125``,
``HydroPopulation`` ``src/cogsworth/hydro/pop.py:195-268``).  :class:`GalacticStageB200` is that
override for the galactic stage (``src/cogsworth/pop.py:1195-1313``): it builds ``w0`` and the event
tables exactly as the reference, flattens everything with one vectorised pass
(``batch.build_orbit_batch``) and makes ONE ``cw_integrate`` call instead of fanning
``integrate_orbit_with_events`` out over a process pool.  ``processes`` is accepted and ignored for
this stage.  Results keep the reference contract: ``orbits`` has ``len(self) + disrupted.sum()``
entries (primaries / bound systems first, then disrupted secondaries, Q7), ``final_pos`` is in kpc
and ``final_vel`` in km/s, failed orbits are dropped with a warning (``pop.py:1269-1311``).

``B200Population`` is ``GalacticStageB200`` mixed into ``cogsworth.pop.Population`` when cogsworth
is importable; here (no gala / COSMIC / astropy in the image) it is mixed into
:class:`TablePopulation`, a minimal holder of the tables the stage reads
(cf. ``EvolvedPopulation``, ``pop.py:2203-2215``).
"""
from __future__ import annotations

import logging
import os
import warnings

import numpy as np

from . import constants as K
from .batch import build_orbit_batch
from .events import draw_sn_angles, identify_event_tables, identify_events
from .kicks import integrate_orbits
from .orbits import OrbitArray
from .potential import MilkyWayPotential

__all__ = ["GalacticStageB200", "TablePopulation", "B200Population", "InitialGalaxy"]


def _val(q, unit):
    if hasattr(q, "unit"):
        import astropy.units as u
        return np.asarray(q.to(u.Unit(unit)).value, dtype=np.float64)
    return np.asarray(q, dtype=np.float64)


class InitialGalaxy:
    """Birth times [Gyr], positions [kpc] and velocities [km/s] (``StarFormationHistory`` stand-in)."""

    def __init__(self, tau, x, y, z, v_x, v_y, v_z):
        self.tau, self.x, self.y, self.z = (np.asarray(a, dtype=np.float64) for a in (tau, x, y, z))
        self.v_x, self.v_y, self.v_z = (np.asarray(a, dtype=np.float64) for a in (v_x, v_y, v_z))

    def __len__(self):
        return len(self.tau)

    def __getitem__(self, mask):
        return InitialGalaxy(self.tau[mask], self.x[mask], self.y[mask], self.z[mask],
                             self.v_x[mask], self.v_y[mask], self.v_z[mask])


class GalacticStageB200:
    """Mixin: overrides the galactic stage of a cogsworth ``Population``."""

    # -- the stage ------------------------------------------------------------------------------
    def perform_galactic_evolution(self, progress_bar=True):
        """Integrate every system (and disrupted secondary) through the Galactic potential on the GPU.

        Same call, same side effects as ``Population.perform_galactic_evolution`` (pop.py:1195);
        ``progress_bar`` is accepted for compatibility (one kernel launch has no per-orbit progress)."""
        self._final_pos = None
        self._final_vel = None
        self._observables = None
        self._escaped = None

        ig = self.initial_galaxy
        pos = [_val(ig.x, "kpc"), _val(ig.y, "kpc"), _val(ig.z, "kpc")]
        vel = [_val(ig.v_x, "km/s"), _val(ig.v_y, "km/s"), _val(ig.v_z, "km/s")]
        tau = _val(ig.tau, "Gyr")

        # per-SN angles are drawn once and persisted (pop.py:1229-1234) so re-runs reproduce
        draw_sn_angles(self.initial_binaries)
        primary_events, secondary_events = identify_event_tables(self)

        retry = {"max_retries": 2, "timestep_multiplier": 0.1}
        retry.update(getattr(self, "orbit_integration_retry_settings", None) or {})
        # The stage runs once per population: its batch and result arrays take page-locked memory only if the library
        # already holds some (pinning 1.4 GB for 10^7 orbits would cost ten times the integration itself); ordinary
        # arrays go through the library's staging ring.
        from ._lib import pinned_policy
        with pinned_policy(getattr(self, "_cw_pinned_policy", "cached")):
            batch = build_orbit_batch(pos, vel, tau, float(_val(self.max_ev_time, "Gyr")),
                                      float(_val(self.timestep_size, "Myr")), self.disrupted,
                                      primary_events, secondary_events)
            orbits, info = integrate_orbits(
                batch, potential=self.galactic_potential, store_all=self.store_entire_orbits,
                integrator=self.integrator, integrator_kwargs=self.integrator_kwargs,
                max_retries=retry["max_retries"], timestep_multiplier=retry["timestep_multiplier"],
                context=getattr(self, "_cw_context", None))
        self._last_integration_info = info

        if orbits.n_failed:
            failed = ~orbits.ok
            orbit_bin_nums = np.concatenate((self.bin_nums, self.bin_nums[self.disrupted]))
            bad_bin_nums = orbit_bin_nums[failed]
            msg = (f"{failed.sum()} orbit(s) failed numerical integration, removing them. This can occur "
                   "due to NaNs in stellar evolution or extreme orbits (e.g. passing directly through the "
                   "galactic centre) that the integrator cannot handle.")
            if getattr(self, "error_file_path", None) is not None and hasattr(self.initial_binaries, "to_hdf"):
                try:
                    file_num, name = 1, os.path.join(self.error_file_path, "failed_integration_binaries.h5")
                    while os.path.exists(name):
                        name = os.path.join(self.error_file_path, f"failed_integration_binaries_{file_num}.h5")
                        file_num += 1
                    self.initial_binaries.loc[bad_bin_nums].to_hdf(name, key="initial_binaries")
                    self.bpp.loc[bad_bin_nums].to_hdf(name, key="bpp")
                    self.kick_info.loc[bad_bin_nums].to_hdf(name, key="kick_info")
                    msg += f" Information for these systems was saved to `{name}`."
                except Exception as exc:      # h5py / pytables missing: keep going, say so
                    msg += f" Could not save them ({type(exc).__name__}: {exc})."
            else:
                msg += " Not saving information for these systems because `error_file_path` is `None`."
            self._warn(msg)
            keep_orbit = ~np.isin(orbit_bin_nums, bad_bin_nums)
            orbits = orbits[keep_orbit]
            new_self = self[~np.isin(self.bin_nums, bad_bin_nums)]
            self.__dict__.update(new_self.__dict__)
        self._orbits = orbits

    # -- accessors that must not loop over Python objects ----------------------------------------
    def _get_final_coords(self):
        """(final_pos [kpc], final_vel [km/s]), shape (len(self) + disrupted.sum(), 3) (pop.py:1315-1349)."""
        orbits = self.orbits
        if isinstance(orbits, OrbitArray):
            fp, fv = orbits.final_coords()
        else:       # a plain object array (e.g. loaded by the reference): reference behaviour
            fp = np.full((len(orbits), 3), np.inf)
            fv = np.full((len(orbits), 3), np.inf)
            for i, o in enumerate(orbits):
                if o is None:
                    warnings.warn("Warning: Detected `None` orbit, entering coordinates as `np.inf`")
                    continue
                fp[i] = _val(o[-1].pos.xyz, "kpc").reshape(3)
                fv[i] = _val(o[-1].vel.d_xyz, "km/s").reshape(3) if hasattr(o[-1].vel.d_xyz, "unit") \
                    else np.asarray(o[-1].vel.d_xyz).reshape(3) * K.KMS_PER_KPC_MYR
        try:
            import astropy.units as u
            fp, fv = fp * u.kpc, fv * u.km / u.s
        except ImportError:
            pass
        self._final_pos, self._final_vel = fp, fv
        return fp, fv

    def save_orbits(self, file_name):
        """The ``orbits`` group of ``Population.save`` (pop.py:1853-1876), straight from the SoA buffers."""
        from .orbit_io import save_orbits
        orbits = self.orbits
        if not isinstance(orbits, OrbitArray):
            raise TypeError("save_orbits writes the OrbitArray produced by perform_galactic_evolution")
        save_orbits(file_name, orbits)

    def load_orbits(self, file_name):
        """Lazy counterpart of the ``orbits`` property's load branch (pop.py:696-719)."""
        from .orbit_io import final_coords_from_file, load_orbits
        self._orbits = load_orbits(file_name)
        self._final_pos, self._final_vel = final_coords_from_file(file_name)
        return self._orbits

    @property
    def primary_orbits(self):
        return self.orbits[:len(self)]

    @property
    def secondary_orbits(self):
        """Orbit of every secondary: the system's own orbit unless disrupted (pop.py:733-742)."""
        n = len(self)
        inds = np.arange(n)
        inds[self.disrupted] = np.arange(n, n + self.disrupted.sum())
        return self.orbits[inds]

    # -- the steps either side of the path: same potential tables, pointwise kernels --------------
    def initial_circular_velocity(self):
        """v_circ [km/s] at every system's birth place AND birth time -- the call
        ``galactic_potential.circular_velocity(q=[x, y, z], t=max_ev_time - tau)`` that draws the birth velocities
        (pop.py:1001-1004); the time matters for gp.TimeInterpolatedPotential."""
        from .kicks import _context
        ig = self.initial_galaxy
        q = np.stack([_val(ig.x, "kpc"), _val(ig.y, "kpc"), _val(ig.z, "kpc")])
        t_birth = (float(_val(self.max_ev_time, "Gyr")) - _val(ig.tau, "Gyr")) * 1000.0            # Myr
        ctx = _context(self.galactic_potential, getattr(self, "_cw_context", None))
        return ctx.circular_velocity(q, t=t_birth) * K.KMS_PER_KPC_MYR

    @property
    def escaped(self):
        """Mask: final speed >= escape speed sqrt(-2 Phi) at the final position (pop.py:874-895), one entry per row of
        ``final_pos``.  As in the reference the potential is evaluated with gala's default time argument (t = 0)."""
        if getattr(self, "_escaped", None) is None:
            from .kicks import _context
            fp = np.asarray(_val(self.final_pos, "kpc"), dtype=np.float64)
            fv = np.asarray(_val(self.final_vel, "km/s"), dtype=np.float64)
            ok = np.isfinite(fp).all(axis=1)
            ctx = _context(self.galactic_potential, getattr(self, "_cw_context", None))
            phi = np.full(len(fp), np.nan)
            if ok.any():
                phi[ok] = ctx.potential_value(np.ascontiguousarray(fp[ok].T), t=0.0)
            with np.errstate(invalid="ignore"):
                v_esc = np.sqrt(-2.0 * phi) * K.KMS_PER_KPC_MYR
                self._escaped = np.sqrt((fv ** 2).sum(axis=1)) >= v_esc
        return self._escaped

    @property
    def final_pos(self):
        if self._final_pos is None:
            self._get_final_coords()
        return self._final_pos

    @property
    def final_vel(self):
        if self._final_vel is None:
            self._get_final_coords()
        return self._final_vel

    @property
    def final_primary_pos(self):
        return self.final_pos[:len(self)]

    @property
    def final_primary_vel(self):
        return self.final_vel[:len(self)]

    def _secondary_inds(self):
        inds = np.arange(len(self))
        inds[self.disrupted] = np.arange(len(self), len(self) + self.disrupted.sum())
        return inds

    @property
    def final_secondary_pos(self):
        return self.final_pos[self._secondary_inds()]

    @property
    def final_secondary_vel(self):
        return self.final_vel[self._secondary_inds()]


class TablePopulation:
    """Minimal stand-in for ``cogsworth.pop.Population`` holding what the galactic stage reads:
    ``initial_galaxy``, ``bpp``, ``kick_info``, ``initial_binaries`` (pandas, indexed by ``bin_num``)
    and the constructor settings of pop.py:121-132 that concern this stage."""

    def __init__(self, initial_galaxy, bpp, kick_info, initial_binaries, galactic_potential=None,
                 max_ev_time=12.0, timestep_size=1.0, store_entire_orbits=True, integrator="dopri853",
                 integrator_kwargs=None, orbit_integration_retry_settings=None, processes=None,
                 error_file_path=None):
        self._initial_galaxy = initial_galaxy
        self._bpp, self._kick_info, self._initial_binaries = bpp, kick_info, initial_binaries
        self.galactic_potential = MilkyWayPotential("v2") if galactic_potential is None else galactic_potential
        self.max_ev_time, self.timestep_size = max_ev_time, timestep_size
        self.store_entire_orbits = store_entire_orbits
        self.integrator, self.integrator_kwargs = integrator, dict(integrator_kwargs or {})
        self.orbit_integration_retry_settings = dict(orbit_integration_retry_settings or {})
        self.processes, self.error_file_path = processes, error_file_path
        self._orbits = self._final_pos = self._final_vel = self._observables = None
        self._disrupted = self._final_bpp = None

    # the attribute surface GalacticStageB200 and events.identify_events use
    initial_galaxy = property(lambda s: s._initial_galaxy)
    bpp = property(lambda s: s._bpp)
    kick_info = property(lambda s: s._kick_info)
    initial_binaries = property(lambda s: s._initial_binaries)
    initC = initial_binaries

    @property
    def bin_nums(self):
        return np.asarray(self._initial_binaries["bin_num"].values)

    @property
    def n_binaries_match(self):
        return len(self._initial_binaries)

    def __len__(self):
        return self.n_binaries_match

    @property
    def final_bpp(self):
        if self._final_bpp is None:
            self._final_bpp = self._bpp.drop_duplicates(subset="bin_num", keep="last")
        return self._final_bpp

    @property
    def disrupted(self):
        """pop.py:861-874: disrupted in kick_info AND final sep < 0 AND an evol_type 11 row."""
        if self._disrupted is None:
            fb, ki, bpp = self.final_bpp, self._kick_info, self._bpp
            self._disrupted = (fb["bin_num"].isin(ki[ki["disrupted"] == 1.0]["bin_num"].unique())
                               & (fb["sep"] < 0.0)
                               & fb["bin_num"].isin(bpp[bpp["evol_type"] == 11.0]["bin_num"])).values
        return self._disrupted

    @property
    def orbits(self):
        if self._orbits is None:
            raise ValueError("No orbits calculated yet, run `perform_galactic_evolution` to do so")
        return self._orbits

    def _warn(self, message):
        logging.getLogger("cogsworth").warning(f"cogsworth warning: {message}")

    def __getitem__(self, mask):
        mask = np.asarray(mask)
        if mask.dtype != bool:
            m = np.zeros(len(self), dtype=bool)
            m[mask] = True
            mask = m
        keep = self.bin_nums[mask]
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        new._initial_galaxy = self._initial_galaxy[mask]
        new._bpp = self._bpp[self._bpp["bin_num"].isin(keep)]
        new._kick_info = self._kick_info[self._kick_info["bin_num"].isin(keep)]
        new._initial_binaries = self._initial_binaries[self._initial_binaries["bin_num"].isin(keep)]
        new._disrupted = new._final_bpp = None
        new._final_pos = new._final_vel = None
        return new


def _make_population_class():
    try:
        from cogsworth.pop import Population as _Ref     # needs gala + COSMIC + astropy
        base = _Ref
    except Exception:
        base = TablePopulation
    return type("B200Population", (GalacticStageB200, base), {
        "__doc__": "cogsworth Population whose galactic stage runs on the B200 kernels "
                   f"(base: {base.__module__}.{base.__name__})."})


B200Population = _make_population_class()
