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
from typing import List, Sequence, Tuple

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
PARAM_STRIDE = 6          # doubles per component handed to cw_set_potential_ex

_COMP_NAMES = {COMP_MIYAMOTO_NAGAI: "MiyamotoNagai", COMP_HERNQUIST: "Hernquist",
               COMP_NFW_SPHERICAL: "NFW", COMP_KEPLER: "Kepler", COMP_PLUMMER: "Plummer",
               COMP_ISOCHRONE: "Isochrone", COMP_JAFFE: "Jaffe", COMP_LOGARITHMIC: "Logarithmic",
               COMP_KUZMIN: "Kuzmin", COMP_BURKERT: "Burkert", COMP_NFW_TRIAXIAL: "TriaxialNFW", COMP_LOG_TRIAXIAL: "TriaxialLogarithmic"}

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

    def describe(self) -> str:
        return ",".join(self.names)

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


def from_gala(potential) -> CompositePotential:
    """Translate a gala potential instance into a parameter table (needs gala; lazy).

    Recognises CompositePotential / MilkyWayPotential made of MiyamotoNagai, MN3ExponentialDisk,
    Hernquist and spherical NFW components; anything else raises NotImplementedError because
    there is deliberately no CPU fallback on this path.
    """
    if isinstance(potential, CompositePotential):
        return potential
    names, types, params = [], [], []
    try:
        items = list(potential.items())
    except AttributeError:
        items = [("potential", potential)]
    G = float(getattr(potential, "G", G_GALACTIC))
    for name, comp in items:
        cls = type(comp).__name__
        p = {k: float(getattr(v, "value", v)) for k, v in comp.parameters.items()}
        if cls == "MiyamotoNagaiPotential":
            names.append(name); types.append(COMP_MIYAMOTO_NAGAI); params.append((p["m"], p["a"], p["b"]))
        elif cls == "MN3ExponentialDiskPotential":
            for i, t in enumerate(mn3_exponential_disk(p["m"], p["h_R"], p["h_z"])):
                names.append(f"{name}.{i}"); types.append(COMP_MIYAMOTO_NAGAI); params.append(t)
        elif cls == "HernquistPotential":
            names.append(name); types.append(COMP_HERNQUIST); params.append((p["m"], p["c"], 0.0))
        elif cls == "NFWPotential" and all(abs(p.get(k, 1.0) - 1.0) < 1e-12 for k in ("a", "b", "c")):
            names.append(name); types.append(COMP_NFW_SPHERICAL); params.append((p["m"], p["r_s"], 0.0))
        elif cls == "NFWPotential":                     # flattened: the spherical formula in the flattened radius
            names.append(name); types.append(COMP_NFW_TRIAXIAL)
            params.append((p["m"], p["r_s"], p.get("a", 1.0), p.get("b", 1.0), p.get("c", 1.0)))
        elif cls == "KeplerPotential":
            names.append(name); types.append(COMP_KEPLER); params.append((p["m"], 0.0, 0.0))
        elif cls == "PlummerPotential":
            names.append(name); types.append(COMP_PLUMMER); params.append((p["m"], p["b"], 0.0))
        elif cls == "IsochronePotential":
            names.append(name); types.append(COMP_ISOCHRONE); params.append((p["m"], p["b"], 0.0))
        elif cls == "KuzminPotential":
            names.append(name); types.append(COMP_KUZMIN); params.append((p["m"], p["a"], 0.0))
        elif cls == "BurkertPotential":
            names.append(name); types.append(COMP_BURKERT); params.append((p["rho"], p["r0"], 0.0))
        elif cls == "NullPotential":
            continue                                   # free motion: contributes nothing
        elif cls == "JaffePotential":
            names.append(name); types.append(COMP_JAFFE); params.append((p["m"], p["c"], 0.0))
        elif (cls == "LogarithmicPotential" and abs(p.get("q1", 1.0) - 1.0) < 1e-12
              and abs(p.get("q2", 1.0) - 1.0) < 1e-12):
            names.append(name); types.append(COMP_LOGARITHMIC); params.append((p["v_c"], p["r_h"], p.get("q3", 1.0)))
        elif cls == "LogarithmicPotential" and abs(p.get("phi", 0.0)) < 1e-15:
            names.append(name); types.append(COMP_LOG_TRIAXIAL)
            params.append((p["v_c"], p["r_h"], p.get("q1", 1.0), p.get("q2", 1.0), p.get("q3", 1.0)))
        else:
            raise NotImplementedError(
                f"component {name!r} ({cls}) has no sm_100a kernel; supported: MiyamotoNagai, "
                "MN3ExponentialDisk, Hernquist, spherical NFW, Kepler, Plummer, Isochrone, Jaffe, Kuzmin, "
                "Burkert, Null, axisymmetric Logarithmic, flattened NFW")
    return CompositePotential(names=names, types=types, params=params, G=G, label=type(potential).__name__)
