#!/usr/bin/env python
"""Pin the oracle (and through it the CUDA path) against REAL gala -- to be run on any machine with gala >= 1.11
and astropy; nothing else of cogsworth's dependencies (COSMIC) is needed.

    python tools/pin_against_gala.py [--kicks-py /path/to/cogsworth/kicks.py] [--out tests/golden/gala_pins.npz]

For every case of tests/_cases.py::gala_pin_cases() it rebuilds the seeded synthetic population (numpy / scipy only),
integrates EVERY orbit with the reference's own `cogsworth.kicks.integrate_orbit_with_events` (imported from an
installed cogsworth, or loaded straight from the given kicks.py -- that file needs only numpy, astropy and gala) in the
matching gala potential, and writes

    <case>_w_final  (6, n)  last sample: kpc, kpc/Myr          <case>_n_nodes  (n,) len(orbit.t)
    <case>_t_last   (n,)    orbit.t[-1] in Myr                  <case>_ev_step  (K,) node index of every kick, computed with
    <case>_failed   (n,)    orbit is None                                    gala's parse_time_specification + the
    <case>_traj_*   full trajectory of the first 4 orbits                    reference's own mask (kicks.py:118-134)

plus `grad_<potential>` / `value_<potential>` at tests/_cases.py::PIN_POINTS and the versions used.  Commit the file:
tests/test_gala_pins.py then holds the oracle (CPU) and the kernels (GPU) to it with BASELINE.json's tolerances
(leapfrog 1e-10, DOPRI853 1e-7 relative, kick nodes exact), which turns every [gala-recall] item of oracle/cw_oracle.c
into a checked one.  This script is test infrastructure; the product never imports gala through it.
"""
import argparse
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_reference_kicks(path):
    try:
        if path is None:
            from cogsworth.kicks import integrate_orbit_with_events
            return integrate_orbit_with_events, "installed cogsworth"
    except Exception as exc:                     # cogsworth's __init__ pulls in COSMIC; fall through to the file
        print(f"import cogsworth failed ({type(exc).__name__}: {exc}); pass --kicks-py", file=sys.stderr)
        raise SystemExit(2)
    spec = importlib.util.spec_from_file_location("cogsworth_kicks_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.integrate_orbit_with_events, os.path.abspath(path)


def gala_potential(key):
    import astropy.units as u
    import gala.potential as gp
    if key == "mw_v1":
        return gp.MilkyWayPotential(version="v1")
    if key == "mw_v2":
        return gp.MilkyWayPotential(version="v2")
    if key == "lm10":
        return gp.LM10Potential()
    if key == "bovy2014":
        return gp.BovyMWPotential2014()
    if key == "evolving_mw":                      # docs/tutorials/potentials/evolving_potentials.ipynb cell 13
        knots = np.linspace(0, 12, 5) * u.Gyr
        f = np.linspace(0.1, 1.0, 5)
        T = gp.TimeInterpolatedPotential
        return gp.CCompositePotential(
            disk=T(gp.MN3ExponentialDiskPotential, knots, m=6.8e10 * u.Msun * f, h_R=2.6 * u.kpc, h_z=0.3 * u.kpc,
                   units="galactic"),
            bulge=T(gp.HernquistPotential, knots, m=5e9 * u.Msun * f, c=1.0 * u.kpc, units="galactic"),
            nucleus=T(gp.HernquistPotential, knots, m=1.81e9 * u.Msun * f, c=0.07 * u.kpc, units="galactic"),
            halo=T(gp.NFWPotential, knots, m=5.4e11 * u.Msun * f, r_s=15.63 * u.kpc, units="galactic"))
    raise KeyError(key)


def kick_nodes(t1, t2, dt, events):
    """Node index of every kick, with the reference's own expressions (kicks.py:97,118-134) on real Quantities."""
    import astropy.units as u
    import gala.integrate as gi
    dt = min(dt, t2 - t1)
    timesteps = gi.parse_time_specification(units=[u.s], t1=t1, t2=t2, dt=dt) * u.s
    if timesteps[-1] < t2:
        timesteps = np.append(timesteps, t2.to(u.s))
    cursor, idx_cursor, out = timesteps[0], 0, []
    for _, ev in events.iterrows():
        mask = (timesteps >= cursor) & (timesteps < (t1 + ev["tphys"] * u.Myr))
        if mask.any():
            idx_cursor = int(np.flatnonzero(mask)[-1])
            cursor = timesteps[idx_cursor]
        else:
            cursor = t1 + ev["tphys"] * u.Myr
        out.append(idx_cursor)
    return out, len(timesteps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kicks-py", default=None, help="path to the reference's src/cogsworth/kicks.py (if cogsworth is not installed)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "gala_pins.npz"))
    ap.add_argument("--cases", default=None, help="comma-separated subset of case names")
    args = ap.parse_args()
    import astropy
    import astropy.units as u
    import gala
    import gala.dynamics as gd
    import gala.integrate as gi
    import pandas as pd
    from _cases import PIN_POINTS, gala_pin_cases, pin_population
    integrate_orbit_with_events, src = load_reference_kicks(args.kicks_py)
    integrators = {"leapfrog": gi.LeapfrogIntegrator, "dop853": gi.DOPRI853Integrator}
    out = {"gala_version": np.array(gala.__version__), "astropy_version": np.array(astropy.__version__),
           "reference_kicks": np.array(src), "points": PIN_POINTS}
    done_pots = set()
    for case in gala_pin_cases():
        name, key, integ, horizon, nb, cfg, dt_myr, kw = case
        if args.cases and name not in args.cases.split(","):
            continue
        pot = gala_potential(key)
        if key not in done_pots:
            done_pots.add(key)
            q = PIN_POINTS * u.kpc
            out[f"grad_{key}"] = pot.gradient(q, t=0.0 * u.Myr).to(u.kpc / u.Myr ** 2).value
            out[f"value_{key}"] = pot.energy(q, t=0.0 * u.Myr).to(u.kpc ** 2 / u.Myr ** 2).value
            out[f"vcirc_{key}"] = pot.circular_velocity(q, t=0.0 * u.Myr).to(u.kpc / u.Myr).value
            if key == "evolving_mw":
                out[f"grad_{key}_t6000"] = pot.gradient(q, t=6000.0 * u.Myr).to(u.kpc / u.Myr ** 2).value
        pop = pin_population(case)
        b = pop.batch
        n, n_bin = b.n_orbits, len(pop.tau_gyr)
        src_of = np.concatenate([np.arange(n_bin), b.secondary_of])
        wf = np.full((6, n), np.nan)
        n_nodes, t_last, failed = np.zeros(n, dtype=np.int64), np.full(n, np.nan), np.zeros(n, dtype=bool)
        ev_step = []
        t2 = pop.max_ev_time_gyr * u.Gyr
        for i in range(n):
            s = src_of[i]
            w0 = gd.PhaseSpacePosition(pos=pop.pos_kpc[:, s] * u.kpc, vel=pop.vel_kms[:, s] * u.km / u.s)
            t1 = t2 - pop.tau_gyr[s] * u.Gyr            # pop.py:1180: max_ev_time - tau
            tab = pop.primary_events if i < n_bin else pop.secondary_events
            rows = np.flatnonzero(tab.owner == s)
            events = None
            if len(rows):
                events = pd.DataFrame({"tphys": tab.tphys[rows], "delta_vsys_x": tab.dv[rows, 0],
                                       "delta_vsys_y": tab.dv[rows, 1], "delta_vsys_z": tab.dv[rows, 2],
                                       "inc": tab.inc[rows], "phase": tab.phase[rows]})
            orbit = integrate_orbit_with_events(w0, t1=t1, t2=t2, dt=dt_myr * u.Myr, potential=pot, events=events,
                                                store_all=True, integrator=integrators[integ],
                                                integrator_kwargs=dict(kw))
            if events is not None:
                ks, _ = kick_nodes(t1, t2, dt_myr * u.Myr, events)
                ev_step.extend(ks)
            if orbit is None:
                failed[i] = True
                continue
            pos = orbit.pos.xyz.to(u.kpc).value.reshape(3, -1)
            vel = orbit.vel.d_xyz.to(u.kpc / u.Myr).value.reshape(3, -1)
            tt = orbit.t.to(u.Myr).value
            wf[:3, i], wf[3:, i] = pos[:, -1], vel[:, -1]
            n_nodes[i], t_last[i] = len(tt), tt[-1]
            if i < 4:
                out[f"{name}_traj_pos_{i}"], out[f"{name}_traj_vel_{i}"], out[f"{name}_traj_t_{i}"] = pos, vel, tt
        out[f"{name}_w_final"], out[f"{name}_n_nodes"], out[f"{name}_t_last"] = wf, n_nodes, t_last
        out[f"{name}_failed"], out[f"{name}_ev_step"] = failed, np.array(ev_step, dtype=np.int64)
        out[f"{name}_w0"] = b.w0                         # guards the generator itself
        print(f"{name}: {n} orbits, {int(failed.sum())} failed, {len(ev_step)} kicks", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
