"""Seeded cases shared by tools/gen_golden_oracle.py (which freezes the oracle's outputs for them) and the tests
that replay them on the CPU oracle and on the CUDA path."""
import numpy as np

from cogsworth_b200 import potential as P

LEAPFROG, DOP853 = 0, 1


def general_cases():
    """(name, potential, integrator, horizon [Gyr] or None = Wagg2022 birth times): the general potentials of SURVEY
    8(f) N4 -- a rotated halo, the incomplete gamma function, per-component time interpolation."""
    base = P.evolving_milky_way([0.0, 12000.0], [1.0, 1.0])
    base = P.CompositePotential(names=base.names, types=base.types, params=base.params)
    f, g = np.array([0.1, 0.2, 0.55, 0.8, 1.0]), np.array([0.3, 0.35, 0.5, 0.9, 1.0])
    per_comp = P.time_interpolated(base, np.linspace(0, 12000, 5),
                                   [p[0] * (g if i == 5 else f) for i, p in enumerate(base.params)])
    return (("gen_lm10_leapfrog", P.LM10Potential(), LEAPFROG, 0.2),
            ("gen_bovy2014_dop853", P.BovyMWPotential2014(), DOP853, 0.2),
            ("gen_evolving_dop853", per_comp, DOP853, None))
