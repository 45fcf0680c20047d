"""Milky Way composite potentials as flat parameter tables.

Host-side mirror of ``gala.potential.MilkyWayPotential(version='v1'|'v2')``, the default
``galactic_potential`` of ``cogsworth.pop.Population`` (reference ``src/cogsworth/pop.py:124``,
``src/cogsworth/kicks.py:53``).  gala itself is an un-vendored dependency, so the parameter
sets are restated here and pinned by the known-answer value printed in
``docs/tutorials/visualisation/gala.ipynb`` (v_circ(8.5 kpc) = 228.89325874 km/s).

A potential is an ordered list of components; the order is the accumulation order of the
gradient (disk, bulge, nucleus, halo -- ``gala.ipynb`` cell 4 repr).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .constants import G_GALACTIC

# component type codes shared with include/cogsworth_b200.h
COMP_MIYAMOTO_NAGAI = 1   # params (m, a, b)
COMP_HERNQUIST = 2        # params (m, c, -)
COMP_NFW_SPHERICAL = 3    # params (m, r_s, -)
# further gala builtins a user can put into a CompositePotential (generic kernel; SURVEY 8f N4,
# reference online-interface/potentials.py:93-131)
COMP_KEPLER = 4           # params (m, -, -)
COMP_PLUMMER = 5          # params (m, b, -)
COMP_ISOCHRONE = 6        # params (m, b, -)
COMP_JAFFE = 7            # params (m, c, -)
COMP_LOGARITHMIC = 8      # params (v_c [kpc/Myr], r_h, q3), q1 = q2 = 1
COMP_KUZMIN = 9           # params (m, a, -)
COMP_BURKERT = 10         # params (rho [Msun/kpc^3], r0, -)
COMP_NFW_TRIAXIAL = 11    # params (m, r_s, a, b, c): gp.NFWPotential with flattening
COMP_LOG_TRIAXIAL = 12    # params (v_c, r_h, q1, q2, q3), phi = 0
# components whose gradient is not (Fx x, Fy y, Fz z): the general kernels (full gradient vector, time argument)
COMP_LONG_MURALI_BAR = 13  # params (m, a, b, c); the bar angle alpha travels as a rotation about z
COMP_LEESUTO_NFW = 14      # params (v_c, r_s, a, b, c): gp.LeeSutoTriaxialNFWPotential
COMP_POWERLAW_CUTOFF = 15  # params (m, alpha, r_c): gp.PowerLawCutoffPotential
PARAM_STRIDE = 6          # doubles per component handed to cw_set_potential_ex
MAX_KNOTS = 64            # time knots of all components together (CW_MAX_KNOTS)

_COMP_NAMES = {COMP_MIYAMOTO_NAGAI: "MiyamotoNagai", COMP_HERNQUIST: "Hernquist",
               COMP_NFW_SPHERICAL: "NFW", COMP_KEPLER: "Kepler", COMP_PLUMMER: "Plummer",
               COMP_ISOCHRONE: "Isochrone", COMP_JAFFE: "Jaffe", COMP_LOGARITHMIC: "Logarithmic",
               COMP_KUZMIN: "Kuzmin", COMP_BURKERT: "Burkert", COMP_NFW_TRIAXIAL: "TriaxialNFW", COMP_LOG_TRIAXIAL: "TriaxialLogarithmic",
               COMP_LONG_MURALI_BAR: "LongMuraliBar", COMP_LEESUTO_NFW: "LeeSutoTriaxialNFW",
               COMP_POWERLAW_CUTOFF: "PowerLawCutoff"}

# Smith+2015 three-Miyamoto-Nagai fit to an exponential disk, positive-density variant
# (the table gala's MN3ExponentialDiskPotential uses with positive_density=True).
_MN3_K_POS = np.array([
    [0.0036, -0.0330, 0.1117, -0.1335, 0.1749],
    [-0.0131, 0.1090, -0.3035, 0.2921, -5.7976],
    [-0.0048, 0.0454, -0.1425, 0.1012, 6.7120],
    [-0.0158, 0.0993, -0.2070, -0.7089, 0.6445],
    [-0.0319, 0.1514, -0.1279, -0.9325, 2.6836],
    [-0.0326, 0.1816, -0.2943, -0.6329, 2.3193],
])


def mn3_exponential_disk(m: float, h_R: float, h_z: float) -> List[Tuple[float, float, float]]:
    """(m_i, a_i, b) of the three Miyamoto-Nagai terms for a sech^2 exponential disk."""
    hzR = h_z / h_R
    b_hR = -0.033 * hzR ** 3 + 0.262 * hzR ** 2 + 0.659 * hzR     # sech2_z=True branch
    x = np.vander([b_hR], N=5)[0]                                # [b^4, b^3, b^2, b, 1]
    param_vec = _MN3_K_POS @ x
    ms = param_vec[:3] * m
    as_ = param_vec[3:] * h_R
    b = b_hR * h_R
    return [(float(ms[i]), float(as_[i]), float(b)) for i in range(3)]


@dataclass
class CompositePotential:
    """Ordered list of analytic components + G, in kpc/Myr/Msun."""
    names: List[str]
    types: List[int]
    params: List[Tuple[float, float, float]]
    G: float = G_GALACTIC
    label: str = "composite"
    units: str = field(default="galactic", repr=False)
    # per component: None, or a 3x3 rotation matrix R (gala's `R=`: the component sees q' = R q)
    rotations: Optional[List[Optional[np.ndarray]]] = None
    # per component: None, or (knot times [Myr], first parameter at the knots): gp.TimeInterpolatedPotential
    knots: Optional[List[Optional[Tuple[np.ndarray, np.ndarray]]]] = None
    interp: str = "cubic"          # "cubic" (natural spline, gala's default) or "linear"
    t_eval: float = 0.0            # time [Myr] of the pointwise evaluations (gradient / value / v_circ)

    @property
    def n_components(self) -> int:
        return len(self.types)

    def type_array(self) -> np.ndarray:
        return np.asarray(self.types, dtype=np.int32)

    def param_array(self) -> np.ndarray:
        """(n, PARAM_STRIDE) table, rows zero-padded (most components have three parameters or fewer)."""
        out = np.zeros((len(self.params), PARAM_STRIDE), dtype=np.float64)
        for i, p in enumerate(self.params):
            out[i, :len(p)] = p
        return out

    def rotation_array(self) -> Optional[np.ndarray]:
        """(n, 9) row-major rotation matrices, NaN rows for unrotated components; None if nothing is rotated."""
        if not self.rotations or all(r is None for r in self.rotations):
            return None
        out = np.full((len(self.types), 9), np.nan, dtype=np.float64)
        for i, r in enumerate(self.rotations):
            if r is not None:
                out[i] = np.asarray(r, dtype=np.float64).reshape(9)
        return out

    def knot_arrays(self):
        """CSR (offsets int32[n+1], times, amplitudes) of the time-interpolated components; None if static."""
        if not self.knots or all(k is None for k in self.knots):
            return None
        off = np.zeros(len(self.types) + 1, dtype=np.int32)
        ts, ys = [], []
        for i, k in enumerate(self.knots):
            if k is not None:
                t, y = np.asarray(k[0], dtype=np.float64).ravel(), np.asarray(k[1], dtype=np.float64).ravel()
                if t.shape != y.shape or t.size == 0:
                    raise ValueError("time knots and amplitudes must have the same non-zero length")
                ts.append(t); ys.append(y)
                off[i + 1] = t.size
        off = np.cumsum(off, dtype=np.int64)
        if off[-1] > MAX_KNOTS:
            raise NotImplementedError(f"more than {MAX_KNOTS} time knots in one potential")
        return off.astype(np.int32), np.ascontiguousarray(np.concatenate(ts)), np.ascontiguousarray(np.concatenate(ys))

    def at_time(self, t_myr: float) -> "CompositePotential":
        """The same potential with the pointwise evaluations (gradient, value, v_circ) taken at time t [Myr]."""
        from dataclasses import replace
        return replace(self, t_eval=float(t_myr))

    def describe(self) -> str:
        return ",".join(self.names)

    def fingerprint(self) -> bytes:
        """Every number a context receives in ``set_potential``: equal fingerprints = the same device tables, so a
        fresh but identical object (``from_gala`` builds one per call) does not trigger a re-upload."""
        import hashlib
        h = hashlib.sha1()
        h.update(np.float64(self.G).tobytes() + self.type_array().tobytes() + self.param_array().tobytes())
        rot, kn = self.rotation_array(), self.knot_arrays()
        h.update(b"R" + (rot.tobytes() if rot is not None else b"-"))
        h.update(b"K" + (b"".join(np.asarray(a).tobytes() for a in kn) if kn is not None else b"-"))
        h.update(self.interp.encode() + np.float64(self.t_eval).tobytes())
        return h.digest()

    def __repr__(self):
        return f"<CompositePotential[{self.label}] {self.describe()}>"


def MilkyWayPotential(version: str = "v2") -> CompositePotential:
    """Parameter table of gala's MilkyWayPotential v1 / MilkyWayPotential2022 (v2)."""
    if version in ("v1", 1, "1"):
        return CompositePotential(
            names=["disk", "bulge", "nucleus", "halo"],
            types=[COMP_MIYAMOTO_NAGAI, COMP_HERNQUIST, COMP_HERNQUIST, COMP_NFW_SPHERICAL],
            params=[(6.8e10, 3.0, 0.28), (5e9, 1.0, 0.0), (1.71e9, 0.07, 0.0), (5.4e11, 15.62, 0.0)],
            label="MilkyWayPotential-v1")
    if version in ("v2", 2, "2", "2022"):
        disk = mn3_exponential_disk(4.7717e10, 2.6, 0.3)
        return CompositePotential(
            names=["disk.0", "disk.1", "disk.2", "bulge", "nucleus", "halo"],
            types=[COMP_MIYAMOTO_NAGAI] * 3 + [COMP_HERNQUIST, COMP_HERNQUIST, COMP_NFW_SPHERICAL],
            params=disk + [(5e9, 1.0, 0.0), (1.8142e9, 0.0688867, 0.0), (5.5427e11, 15.626, 0.0)],
            label="MilkyWayPotential2022")
    raise ValueError(f"unknown MilkyWayPotential version {version!r}")


def MilkyWayPotential2022() -> CompositePotential:
    return MilkyWayPotential("v2")


def rotation_z(angle: float) -> np.ndarray:
    """R with x' = x cos(angle) + y sin(angle), y' = -x sin(angle) + y cos(angle): the frame of a component whose
    major axis is turned by `angle` from the x axis (LogarithmicPotential's phi -- Law & Majewski 2010 eq. 3's
    C1..C3 -- and LongMuraliBarPotential's alpha) [gala-recall]."""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


def LM10Potential() -> CompositePotential:
    """gala's LM10Potential (Law & Majewski 2010; reference online-interface/potentials.py:101) [gala-recall]:
    MN disk + Hernquist bulge + triaxial logarithmic halo turned by phi = 97 deg."""
    v_c = np.sqrt(2.0) * 121.858 / 977.7922216807891            # km/s -> kpc/Myr
    return CompositePotential(
        names=["disk", "bulge", "halo"],
        types=[COMP_MIYAMOTO_NAGAI, COMP_HERNQUIST, COMP_LOG_TRIAXIAL],
        params=[(1e11, 6.5, 0.26), (3.4e10, 0.7, 0.0), (v_c, 12.0, 1.38, 1.0, 1.36)],
        rotations=[None, None, rotation_z(np.deg2rad(97.0))], label="LM10Potential")


def BovyMWPotential2014() -> CompositePotential:
    """gala's BovyMWPotential2014 (reference online-interface/potentials.py:99) [gala-recall]: MN disk +
    power-law-cutoff bulge + NFW halo."""
    return CompositePotential(
        names=["disk", "bulge", "halo"],
        types=[COMP_MIYAMOTO_NAGAI, COMP_POWERLAW_CUTOFF, COMP_NFW_SPHERICAL],
        params=[(68193902782.346756, 3.0, 0.28), (4501365375.06545, 1.8, 1.9), (4.3683325e11, 16.0, 0.0)],
        label="BovyMWPotential2014")


def LongMuraliBarPotential(m, a, b, c, alpha=0.0) -> CompositePotential:
    """gp.LongMuraliBarPotential (reference online-interface/potentials.py:130)."""
    return CompositePotential(names=["bar"], types=[COMP_LONG_MURALI_BAR], params=[(m, a, b, c)],
                              rotations=[rotation_z(alpha) if alpha != 0.0 else None], label="LongMuraliBarPotential")


def LeeSutoTriaxialNFWPotential(v_c, r_s, a=1.0, b=0.9, c=0.8) -> CompositePotential:
    """gp.LeeSutoTriaxialNFWPotential (reference online-interface/potentials.py:128)."""
    return CompositePotential(names=["halo"], types=[COMP_LEESUTO_NFW], params=[(v_c, r_s, a, b, c)],
                              label="LeeSutoTriaxialNFWPotential")


def time_interpolated(base: CompositePotential, time_knots_myr, amplitudes, interp: str = "cubic") -> CompositePotential:
    """gp.TimeInterpolatedPotential(cls, time_knots, m=masses, ...) for every component of `base`
    (docs/tutorials/potentials/evolving_potentials.ipynb cells 5 and 13): the first parameter (mass / density) of
    component i follows `amplitudes[i]` (array over the knots; a scalar or None keeps it static).  An
    MN3ExponentialDisk's three terms scale together: pass the same mass FRACTIONS times each term's mass."""
    from dataclasses import replace
    t = np.asarray(time_knots_myr, dtype=np.float64)
    if len(amplitudes) != base.n_components:
        raise ValueError("one amplitude entry per component")
    knots = []
    for a in amplitudes:
        if a is None or np.ndim(a) == 0:
            knots.append(None)
        else:
            knots.append((t, np.asarray(a, dtype=np.float64)))
    params = [tuple(p) for p in base.params]
    for i, a in enumerate(amplitudes):
        if a is not None and np.ndim(a) == 0:
            params[i] = (float(a),) + params[i][1:]
    return replace(base, params=params, knots=knots, interp=interp, label=base.label + "[time-interpolated]")


def evolving_milky_way(time_knots_myr, mass_fractions, interp: str = "cubic") -> CompositePotential:
    """The tutorial's evolving Milky Way (evolving_potentials.ipynb cell 13): MN3 disk 6.8e10, Hernquist bulge 5e9 /
    nucleus 1.81e9 (c = 0.07), NFW halo 5.4e11 (r_s = 15.63), every mass scaled by mass_fractions(t)."""
    f = np.asarray(mass_fractions, dtype=np.float64)
    disk = mn3_exponential_disk(6.8e10, 2.6, 0.3)
    base = CompositePotential(
        names=["disk.0", "disk.1", "disk.2", "bulge", "nucleus", "halo"],
        types=[COMP_MIYAMOTO_NAGAI] * 3 + [COMP_HERNQUIST, COMP_HERNQUIST, COMP_NFW_SPHERICAL],
        params=disk + [(5e9, 1.0, 0.0), (1.81e9, 0.07, 0.0), (5.4e11, 15.63, 0.0)], label="evolving_mw")
    return time_interpolated(base, time_knots_myr, [p[0] * f for p in base.params], interp)


def _plain(v, unit=None):
    """float / ndarray of a number, ndarray or astropy Quantity (converted to `unit` when it is one)."""
    if unit is not None and hasattr(v, "to_value"):
        try:
            v = v.to_value(unit)
        except Exception:
            v = getattr(v, "value", v)
    else:
        v = getattr(v, "value", v)
    return np.asarray(v, dtype=np.float64)


def _galactic_value(v, units):
    """A parameter as a plain number in gala's ``galactic`` unit system (kpc, Myr, Msun, rad).  gala stores parameters
    as Quantities in the units the user gave (LM10's v_c in km/s, a scale radius in pc, ...): they are DECOMPOSED into
    the potential's unit system, not read by ``.value``."""
    if hasattr(v, "decompose") and units is not None:
        try:
            return np.asarray(v.decompose(units).value, dtype=np.float64)
        except Exception:
            pass
    return _plain(v)


def _require_galactic(units):
    """Everything below (G, the Myr time grid, kpc / kpc/Myr state) assumes the galactic unit system."""
    if units is None:
        return
    try:
        import astropy.units as u
        scale = [float((1.0 * q).decompose(units).value) for q in (u.kpc, u.Myr, u.Msun)]
    except Exception:
        return                                    # stand-ins without astropy: nothing to check
    if not np.allclose(scale, 1.0, rtol=1e-12, atol=0):
        raise NotImplementedError("cogsworth_b200 integrates in gala's galactic unit system (kpc, Myr, Msun); this "
                                  f"potential uses {units!r}")


def from_gala(potential) -> CompositePotential:
    """Translate a gala potential instance into a parameter table (needs gala; lazy).

    Recognises CompositePotential / MilkyWayPotential / LM10Potential / BovyMWPotential2014 made of the components
    listed in the error message below, each optionally rotated (``R=``, ``phi``, ``alpha``) or wrapped in a
    ``gp.TimeInterpolatedPotential`` whose only time-dependent parameter is the mass; anything else raises
    NotImplementedError because there is deliberately no CPU fallback on this path.
    """
    if isinstance(potential, CompositePotential):
        return potential
    names, types, params, rotations, knots = [], [], [], [], []
    interp = "cubic"
    try:
        items = list(potential.items())
    except AttributeError:
        items = [("potential", potential)]
    _require_galactic(getattr(potential, "units", None))
    G = float(_plain(getattr(potential, "G", G_GALACTIC)))

    def add(name, t, p, R=None, kn=None):
        names.append(name); types.append(t); params.append(tuple(float(x) for x in p))
        rotations.append(R); knots.append(kn)

    for name, comp in items:
        cls = type(comp).__name__
        t_knots = None
        if cls == "TimeInterpolatedPotential":        # gala >= 1.10: (potential_cls, time_knots, **parameter arrays)
            inner = getattr(comp, "potential_cls", None) or getattr(comp, "_potential_cls", None)
            tk = getattr(comp, "time_knots", None)
            if tk is None:
                tk = getattr(comp, "_time_knots", None)
            if inner is None or tk is None:
                raise NotImplementedError("TimeInterpolatedPotential: cannot read potential_cls / time_knots of this gala version")
            cls = inner.__name__
            t_knots = _plain(tk, "Myr").ravel()
            interp = str(getattr(comp, "interpolation_method", getattr(comp, "_interpolation_method", "cubic")))
            if interp not in ("cubic", "linear"):
                raise NotImplementedError(f"TimeInterpolatedPotential interpolation {interp!r} (cubic and linear are built)")
        cunits = getattr(comp, "units", None) or getattr(potential, "units", None)
        praw = {k: _galactic_value(v, cunits) for k, v in comp.parameters.items()}
        varying = [k for k, v in praw.items() if v.ndim > 0 and v.size > 1 and not np.all(v == v.flat[0])]
        if t_knots is None and varying:
            raise NotImplementedError(f"component {name!r}: array-valued parameters {varying}")
        if [k for k in varying if k not in ("m", "rho")]:
            raise NotImplementedError(f"component {name!r}: only a time-dependent mass is built, not {varying}")
        p = {k: float(v.flat[0]) for k, v in praw.items() if v.size}
        amp_t = None
        for k in ("m", "rho"):
            if k in varying:
                amp_t = praw[k].ravel()
        R = getattr(comp, "R", None)
        if R is not None:
            R = _plain(R)
            if R.shape != (3, 3):
                raise NotImplementedError(f"component {name!r}: time-dependent rotation matrices are not built")
            if np.allclose(R, np.eye(3), rtol=0, atol=0):
                R = None
        kn = (lambda a: (t_knots, a)) if amp_t is not None else (lambda a: None)
        if cls == "MiyamotoNagaiPotential":
            add(name, COMP_MIYAMOTO_NAGAI, (p["m"], p["a"], p["b"]), R, kn(amp_t))
        elif cls == "MN3ExponentialDiskPotential":
            for i, t in enumerate(mn3_exponential_disk(p["m"], p["h_R"], p["h_z"])):
                add(f"{name}.{i}", COMP_MIYAMOTO_NAGAI, t, R, kn(None if amp_t is None else amp_t * (t[0] / p["m"])))
        elif cls == "HernquistPotential":
            add(name, COMP_HERNQUIST, (p["m"], p["c"], 0.0), R, kn(amp_t))
        elif cls == "NFWPotential" and all(abs(p.get(k, 1.0) - 1.0) < 1e-12 for k in ("a", "b", "c")):
            add(name, COMP_NFW_SPHERICAL, (p["m"], p["r_s"], 0.0), R, kn(amp_t))
        elif cls == "NFWPotential":                     # flattened: the spherical formula in the flattened radius
            add(name, COMP_NFW_TRIAXIAL, (p["m"], p["r_s"], p.get("a", 1.0), p.get("b", 1.0), p.get("c", 1.0)), R, kn(amp_t))
        elif cls == "KeplerPotential":
            add(name, COMP_KEPLER, (p["m"], 0.0, 0.0), R, kn(amp_t))
        elif cls == "PlummerPotential":
            add(name, COMP_PLUMMER, (p["m"], p["b"], 0.0), R, kn(amp_t))
        elif cls == "IsochronePotential":
            add(name, COMP_ISOCHRONE, (p["m"], p["b"], 0.0), R, kn(amp_t))
        elif cls == "KuzminPotential":
            add(name, COMP_KUZMIN, (p["m"], p["a"], 0.0), R, kn(amp_t))
        elif cls == "BurkertPotential":
            add(name, COMP_BURKERT, (p["rho"], p["r0"], 0.0), R, kn(amp_t))
        elif cls == "NullPotential":
            continue                                   # free motion: contributes nothing
        elif cls == "JaffePotential":
            add(name, COMP_JAFFE, (p["m"], p["c"], 0.0), R, kn(amp_t))
        elif cls == "PowerLawCutoffPotential":
            add(name, COMP_POWERLAW_CUTOFF, (p["m"], p["alpha"], p["r_c"]), R, kn(amp_t))
        elif cls == "LongMuraliBarPotential":
            Ra = rotation_z(p["alpha"]) if abs(p.get("alpha", 0.0)) > 0.0 else None
            if Ra is not None and R is not None:
                Ra = Ra @ R
            add(name, COMP_LONG_MURALI_BAR, (p["m"], p["a"], p["b"], p["c"]), Ra if Ra is not None else R, kn(amp_t))
        elif cls == "LeeSutoTriaxialNFWPotential":
            add(name, COMP_LEESUTO_NFW, (p["v_c"], p["r_s"], p.get("a", 1.0), p.get("b", 1.0), p.get("c", 1.0)), R)
        elif (cls == "LogarithmicPotential" and abs(p.get("q1", 1.0) - 1.0) < 1e-12
              and abs(p.get("q2", 1.0) - 1.0) < 1e-12 and abs(p.get("phi", 0.0)) < 1e-15):
            add(name, COMP_LOGARITHMIC, (p["v_c"], p["r_h"], p.get("q3", 1.0)), R)
        elif cls == "LogarithmicPotential":
            Rp = rotation_z(p["phi"]) if abs(p.get("phi", 0.0)) >= 1e-15 else None
            if Rp is not None and R is not None:
                Rp = Rp @ R
            add(name, COMP_LOG_TRIAXIAL, (p["v_c"], p["r_h"], p.get("q1", 1.0), p.get("q2", 1.0), p.get("q3", 1.0)),
                Rp if Rp is not None else R)
        else:
            raise NotImplementedError(
                f"component {name!r} ({cls}) has no sm_100a kernel; supported: MiyamotoNagai, "
                "MN3ExponentialDisk, Hernquist, NFW (spherical / flattened), Kepler, Plummer, Isochrone, Jaffe, Kuzmin, "
                "Burkert, Null, Logarithmic, PowerLawCutoff, LongMuraliBar, LeeSutoTriaxialNFW, each optionally "
                "rotated or mass-interpolated in time (TimeInterpolatedPotential)")
    return CompositePotential(names=names, types=types, params=params, G=G, label=type(potential).__name__,
                              rotations=rotations if any(r is not None for r in rotations) else None,
                              knots=knots if any(k is not None for k in knots) else None, interp=interp)
