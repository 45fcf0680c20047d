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


def rel_diff(a, b):
    """Per-orbit worst relative difference over the six coordinates, scaled by the orbit's own size
    (|pos| for positions, |vel| for velocities) so a coordinate crossing zero does not blow it up."""
    sp = np.maximum(np.linalg.norm(b[:3], axis=0), 1e-3)
    sv = np.maximum(np.linalg.norm(b[3:], axis=0), 1e-5)
    return np.maximum(np.abs(a[:3] - b[:3]).max(axis=0) / sp, np.abs(a[3:] - b[3:]).max(axis=0) / sv)


def oracle_run(O, pot, b, integ, **kw):
    return O.integrate_batch(pot, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                             integrator=integ, **kw)


def sensitive_orbits(O, pot, b, integ, ref=None, threshold=1e-11, seeds=(1, 2, 3), eps=2.0 ** -52, **kw):
    """Which orbits amplify LAST-BIT differences of the gradient beyond ``threshold``: the oracle is re-run with every
    gradient component multiplied by 1 + eps * {-1, 0, 1} (``oracle.set_rounding_noise``) -- what any equivalent but
    differently rounded implementation (another libm, FMA contraction, this repo's CUDA kernels, gala on another
    machine) does to the arithmetic -- and an orbit is SENSITIVE if any of the noisy runs leaves the clean one by more
    than ``threshold``.  No two implementations can agree on those to the parity tolerance; every other orbit must.
    Returns (mask, worst noisy-vs-clean difference per orbit, clean result)."""
    ref = oracle_run(O, pot, b, integ, **kw) if ref is None else ref
    worst = np.zeros(b.n_orbits)
    try:
        for seed in seeds:
            O.set_rounding_noise(eps, seed)
            r = oracle_run(O, pot, b, integ, **kw)
            worst = np.maximum(worst, np.nan_to_num(rel_diff(r["w_final"], ref["w_final"]), nan=np.inf))
    finally:
        O.set_rounding_noise(0.0)
    return worst > threshold, worst, ref


def assert_parity_outside_the_sensitive_set(d, sens, tol, max_sensitive_fraction, what=""):
    """The parity statement of BASELINE.json's north star, made checkable: EVERY orbit that does not amplify last-bit
    noise agrees within ``tol``; equivalently, every orbit beyond ``tol`` is one the oracle cannot reproduce against
    its own rounding noise.  The sensitive share itself is bounded so the statement cannot become vacuous."""
    out = d >= tol
    bad = out & ~sens
    assert not bad.any(), (f"{what}: {int(bad.sum())} orbit(s) beyond {tol:g} that are NOT sensitive to rounding noise "
                           f"(worst {d[bad].max():.2e}); outliers {int(out.sum())}, sensitive {int(sens.sum())} of {len(d)}")
    assert sens.mean() <= max_sensitive_fraction, (what, float(sens.mean()))
    ok = ~sens
    return dict(n=len(d), sensitive=int(sens.sum()), outliers=int(out.sum()),
                max_regular=float(d[ok].max()) if ok.any() else 0.0, median=float(np.median(d)))


# ---- cases pinned against REAL gala by tools/pin_against_gala.py (tests/test_gala_pins.py consumes the file) -------
def gala_pin_cases():
    """(name, potential key, integrator name, horizon spec, n_binaries, cfg_id, dt_myr, integrator_kwargs).

    Each case is one seeded synthetic population (cogsworth_b200.synthetic.make_population -- numpy / scipy only, so it
    is reproducible on the gala machine) integrated orbit by orbit with cogsworth.kicks.integrate_orbit_with_events.
    Together they cover every [gala-recall] item of the oracle: the leapfrog recipe, the un-kicked end point (Q1), the
    short appended interval (Q2), restart-vs-dense DOPRI853, and the general potentials' conventions."""
    return (
        ("cfg1_leapfrog_v1", "mw_v1", "leapfrog", 0.2, 96, 1, 1.0, {}),
        ("cfg4_leapfrog_v2", "mw_v2", "leapfrog", 1.0, 96, 4, 1.0, {}),
        ("cfg2_leapfrog_v2_wagg", "mw_v2", "leapfrog", None, 24, 2, 1.0, {}),
        ("leapfrog_v2_odd_dt", "mw_v2", "leapfrog", (0.0002, 0.3), 64, 31, 0.7, {}),       # Q1, Q2, clipped dt
        ("cfg1_dop853_v2", "mw_v2", "dop853", 0.2, 96, 1, 1.0, {}),
        ("cfg5_dop853_v2_1e-8", "mw_v2", "dop853", 1.0, 48, 5, 1.0, {"atol": 1e-8, "rtol": 1e-8}),
        ("lm10_leapfrog", "lm10", "leapfrog", 0.2, 48, 1, 1.0, {}),
        ("bovy2014_dop853", "bovy2014", "dop853", 0.2, 48, 1, 1.0, {}),
        ("evolving_mw_dop853", "evolving_mw", "dop853", None, 24, 1, 1.0, {}),
    )


def pin_potential(key):
    """This package's table for a pin-case potential key (tools/pin_against_gala.py builds the gala object)."""
    if key == "mw_v1":
        return P.MilkyWayPotential("v1")
    if key == "mw_v2":
        return P.MilkyWayPotential("v2")
    if key == "lm10":
        return P.LM10Potential()
    if key == "bovy2014":
        return P.BovyMWPotential2014()
    if key == "evolving_mw":
        return P.evolving_milky_way(np.linspace(0.0, 12000.0, 5), np.linspace(0.1, 1.0, 5))
    raise KeyError(key)


def pin_population(case):
    """The seeded population of a pin case (birth velocities are always drawn in the static MilkyWayPotential v2)."""
    from cogsworth_b200.synthetic import make_population
    name, key, integ, horizon, nb, cfg, dt_myr, kw = case
    draw_in = P.MilkyWayPotential("v1") if key == "mw_v1" else P.MilkyWayPotential("v2")
    pop = make_population(nb, cfg_id=cfg, potential=draw_in, horizon_gyr=horizon, timestep_myr=dt_myr)
    pop.potential = pin_potential(key)
    return pop


#: points [kpc] at which the pin file also holds every potential's gradient / value (conventions of rotated components)
PIN_POINTS = np.array([[8.5, 0.0, 0.0], [3.2, -4.1, 0.7], [-0.3, 0.2, 0.05], [15.0, 9.0, -20.0], [0.9, 2.2, -1.4],
                       [-30.0, 2.0, 11.0]]).T
