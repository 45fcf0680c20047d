"""Synthetic populations of the BASELINE.json configs: INPUT GENERATION ONLY.

Stands in for the two host stages that sit upstream of the hot path and are outside its scope:
the Wagg2022 star-formation-history sampler (reference ``src/cogsworth/sfh.py:1026-1165``; birth
time, position) with the birth-velocity rule of ``Population.sample_initial_galaxy``
(``src/cogsworth/pop.py:1001-1013``), and COSMIC's ``bpp`` / ``kick_info`` supernova rows
(drawn from simple distributions; COSMIC is Fortran and excluded by the north star).
Everything is seeded: ``numpy.random.default_rng(20221101 + cfg_id)``; draws happen in the order
ICs, then events.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.integrate import quad
from scipy.special import lambertw
from scipy.stats import beta as beta_dist

from . import constants as K
from .batch import EventTable, OrbitBatch, build_orbit_batch
from .potential import CompositePotential, MilkyWayPotential

SEED0 = 20221101


def _host_circular_velocity(pot: CompositePotential, x, y, z):
    """v_circ = sqrt(r |dPhi/dr|) in kpc/Myr, vectorised numpy -- used ONLY to draw birth velocities
    for synthetic inputs (upstream of the hot path; the product's own version is cw_circular_velocity)."""
    r = np.sqrt(x * x + y * y + z * z)
    qdotg = np.zeros_like(r)
    if any(t > 12 for t in pot.types) or pot.rotation_array() is not None or pot.knot_arrays() is not None:
        raise NotImplementedError("synthetic birth velocities need a static, unrotated potential of the diagonal "
                                  "components: draw them in one and integrate in the other")
    for t, prm in zip(pot.types, pot.params):
        m, a, b = (tuple(prm) + (0.0, 0.0))[:3]
        GM = pot.G * m
        if t == 1:
            sz = np.sqrt(z * z + b * b)
            zd = a + sz
            f = GM * (x * x + y * y + zd * zd) ** -1.5
            qdotg += f * (x * x + y * y) + f * z * z * (1.0 + a / sz)
        elif t == 2:
            qdotg += GM / ((r + a) ** 2 * r) * r * r
        elif t == 4:
            qdotg += GM / r
        elif t == 5:
            qdotg += GM * (r * r + a * a) ** -1.5 * r * r
        elif t == 6:
            s = np.sqrt(r * r + a * a)
            qdotg += GM / (s * (s + a) ** 2) * r * r
        elif t == 7:
            qdotg += GM / (r + a)
        elif t == 11:                                    # flattened NFW (m, r_s, a, b, c)
            qa, qb, qc = prm[2], prm[3], prm[4]
            u = np.sqrt(x * x / qa ** 2 + y * y / qb ** 2 + z * z / qc ** 2) / a
            fac = GM / a / u ** 3 / (a * a) * (np.log(1 + u) - u / (1 + u))
            qdotg += fac * (x * x / qa ** 2 + y * y / qb ** 2 + z * z / qc ** 2)
        elif t == 12:                                    # triaxial logarithmic (v_c, r_h, q1, q2, q3)
            s2 = x * x / prm[2] ** 2 + y * y / prm[3] ** 2 + z * z / prm[4] ** 2
            qdotg += m * m * s2 / (a * a + s2)
        elif t == 9:
            az = a + np.abs(z)
            f = GM * (x * x + y * y + az * az) ** -1.5
            qdotg += f * (x * x + y * y) + f * az * np.abs(z)
        elif t == 10:                                    # Burkert (rho, r0)
            xb = r / a
            qdotg += np.pi * pot.G * m * a / xb ** 2 * (np.log(1 + xb * xb) + 2 * np.log(1 + xb) - 2 * np.arctan(xb)) * r
        elif t == 8:                                     # (v_c, r_h, q3)
            fac = m * m / (a * a + x * x + y * y + z * z / (b * b))
            qdotg += fac * (x * x + y * y) + fac * z * z / (b * b)
        else:
            u = r / a
            qdotg += GM / a / (u ** 3) / (a * a) * (np.log(1 + u) - u / (1 + u)) * r * r
    return np.sqrt(np.abs(qdotg))


def wagg2022_initial_conditions(n, rng, potential=None, galaxy_age=12.0, tsfr=6.8, alpha=0.3,
                                v_dispersion_kms=5.0):
    """Birth (tau [Gyr], pos (3,n) [kpc], vel (3,n) [km/s]) of ``n`` systems, Wagg+2022-like."""
    potential = MilkyWayPotential("v2") if potential is None else potential
    comp = rng.choice(3, size=n, p=np.array([2.585, 2.585, 0.91]) / (2 * 2.585 + 0.91))
    U = rng.random(n)
    tau = np.empty(n)
    # low-alpha disc: 0-8 Gyr (sfh.py:1050-1052); high-alpha: 8-12 Gyr (sfh.py:1083-1086)
    norm_lo = 1 / quad(lambda x: np.exp(-(galaxy_age - x) / tsfr), 0, 8)[0]
    norm_hi = 1 / quad(lambda x: np.exp(-(galaxy_age - x) / tsfr), 8, 12)[0]
    lo, hi, bu = comp == 0, comp == 1, comp == 2
    tau[lo] = tsfr * np.log(U[lo] * np.exp(galaxy_age / tsfr) / (norm_lo * tsfr) + 1)
    tau[hi] = tsfr * np.log(U[hi] * np.exp(galaxy_age / tsfr) / (norm_hi * tsfr) + np.exp(8 / tsfr))
    tau[bu] = beta_dist.rvs(a=2, b=3, loc=6, scale=6, size=int(bu.sum()), random_state=rng)  # sfh.py:1115
    # radii: inverse CDF of an exponential disc (sfh.py:1003); scale lengths sfh.py:1055,1065,1096
    R_d = np.where(lo, 4.0 * (1 - alpha * tau / 8.0), np.where(hi, 1 / 0.43, 1.5))
    rho = -R_d * (lambertw((rng.random(n) - 1) / np.e, k=-1).real + 1)
    h_z = np.where(lo, 0.3, np.where(hi, 0.95, 0.2))
    z = h_z * np.log(1 - rng.random(n)) * rng.choice([-1.0, 1.0], size=n)
    phi = rng.uniform(0, 2 * np.pi, n)
    x, y = rho * np.sin(phi), rho * np.cos(phi)                  # sfh.py:425-426
    # velocities (pop.py:1001-1013): N((0, v_circ, 0), sigma/sqrt(3)) in (v_R, v_T, v_z)
    v_circ = _host_circular_velocity(potential, x, y, z) * K.KMS_PER_KPC_MYR
    sig = v_dispersion_kms / np.sqrt(3)
    v_R, v_T, v_zz = rng.normal(0, sig, n), rng.normal(v_circ, sig), rng.normal(0, sig, n)
    ang = np.arctan2(y, x)
    v_x = v_R * np.cos(ang) - v_T * np.sin(ang)                  # sfh.py:327-345
    v_y = v_R * np.sin(ang) + v_T * np.cos(ang)
    return tau, np.stack([x, y, z]), np.stack([v_x, v_y, v_zz])


def synthetic_supernovae(n, tau_gyr, rng, f_sn=1.0, p_second=0.3, p_disrupt=0.8, sigma1=50.0, sigma2=150.0,
                         tphys_range=(3.0, 50.0)):
    """Stand-in for COSMIC's SN rows: returns (primary EventTable, secondary EventTable, disrupted mask)
    built with exactly the rules of ``events.identify_events`` (events.py:47-70)."""
    kicked = rng.random(n) < f_sn
    t_sn1 = rng.uniform(*tphys_range, n)
    has2 = kicked & (rng.random(n) < p_second)
    t_sn2 = t_sn1 + rng.uniform(0.1, 20.0, n)
    horizon = tau_gyr * 1000.0
    ok1 = kicked & (t_sn1 < horizon)                 # events later than the evolution time are discarded
    ok2 = has2 & ok1 & (t_sn2 < horizon)
    n_sn = ok1.astype(int) + ok2.astype(int)
    disrupt_draw = rng.random(n) < p_disrupt
    disrupted = (n_sn > 0) & disrupt_draw            # disruption happens at the LAST supernova
    dv1_a, dv1_b = rng.normal(0, sigma1, (n, 3)), rng.normal(0, sigma1, (n, 3))
    dv2_a, dv2_b = rng.normal(0, sigma2, (n, 3)), rng.normal(0, sigma2, (n, 3))
    phase = rng.uniform(0, 2 * np.pi, (n, 2))
    inc = np.arccos(2 * rng.random((n, 2)) - 1.0)

    i1, i2 = np.flatnonzero(ok1), np.flatnonzero(ok2)
    owner = np.concatenate([i1, i2])
    order = np.lexsort((np.concatenate([np.zeros(len(i1)), np.ones(len(i2))]), owner))
    owner = owner[order]
    which = np.concatenate([np.zeros(len(i1), dtype=int), np.ones(len(i2), dtype=int)])[order]
    tphys = np.where(which == 0, t_sn1[owner], t_sn2[owner])
    dv1 = np.where((which == 0)[:, None], dv1_a[owner], dv1_b[owner])
    dv2 = np.where((which == 0)[:, None], dv2_a[owner], dv2_b[owner])
    ph, ic = phase[owner, which], inc[owner, which]
    is_last = which == (n_sn[owner] - 1)
    row_disrupted = disrupted[owner] & is_last       # kick_info.disrupted of that SN row
    prim = EventTable(owner, tphys, dv1, ph, ic)
    keep = disrupted[owner]
    dv_sec = np.where(row_disrupted[:, None], dv2, dv1)
    sec = EventTable(owner[keep], tphys[keep], dv_sec[keep], ph[keep], ic[keep])
    return prim, sec, disrupted


@dataclass
class SyntheticPopulation:
    batch: OrbitBatch
    tau_gyr: np.ndarray
    pos_kpc: np.ndarray
    vel_kms: np.ndarray
    disrupted: np.ndarray
    primary_events: EventTable
    secondary_events: EventTable
    potential: CompositePotential
    max_ev_time_gyr: float
    timestep_myr: float


def make_population(n_binaries, cfg_id=0, potential=None, f_sn=1.0, horizon_gyr=None, max_ev_time_gyr=12.0,
                    timestep_myr=1.0, plunging_fraction=0.0, seed=None) -> SyntheticPopulation:
    """A seeded synthetic population + its flattened OrbitBatch.

    horizon_gyr : None -> every system keeps its Wagg2022 birth time (variable step counts);
    a number -> every orbit is integrated for exactly that long (configs 1, 3, 4, 5).
    plunging_fraction : share of systems put on nearly radial orbits (|L_z| < 50 kpc km/s), config 5.
    """
    potential = MilkyWayPotential("v2") if potential is None else potential
    rng = np.random.default_rng(SEED0 + cfg_id if seed is None else seed)
    tau, pos, vel = wagg2022_initial_conditions(n_binaries, rng, potential, galaxy_age=max_ev_time_gyr)
    if isinstance(horizon_gyr, (tuple, list)):       # every system its own horizon, uniform in [lo, hi] Gyr
        tau = rng.uniform(float(horizon_gyr[0]), float(horizon_gyr[1]), n_binaries)
    elif horizon_gyr is not None:
        tau = np.full(n_binaries, float(horizon_gyr))
    if plunging_fraction > 0:
        k = rng.random(n_binaries) < plunging_fraction
        R = np.sqrt(pos[0] ** 2 + pos[1] ** 2)
        lz_target = rng.uniform(-50.0, 50.0, n_binaries)            # kpc km/s
        vT = lz_target / np.maximum(R, 1e-3)
        ang = np.arctan2(pos[1], pos[0])
        vR = rng.normal(0, 30.0, n_binaries)
        vel[0] = np.where(k, vR * np.cos(ang) - vT * np.sin(ang), vel[0])
        vel[1] = np.where(k, vR * np.sin(ang) + vT * np.cos(ang), vel[1])
    if f_sn > 0:
        prim, sec, disrupted = synthetic_supernovae(n_binaries, tau, rng, f_sn=f_sn)
    else:
        prim, sec, disrupted = EventTable.empty(), EventTable.empty(), np.zeros(n_binaries, dtype=bool)
    batch = build_orbit_batch(pos, vel, tau, max_ev_time_gyr, timestep_myr, disrupted, prim, sec)
    return SyntheticPopulation(batch, tau, pos, vel, disrupted, prim, sec, potential, max_ev_time_gyr, timestep_myr)


def population_with_n_orbits(n_orbits, **kw) -> SyntheticPopulation:
    """Config 4 / 5 are specified in ORBITS (primaries + disrupted secondaries): pick the number of
    binaries whose expected orbit count matches, then trim the disrupted tail to hit it exactly."""
    f_sn = kw.get("f_sn", 1.0)
    exp_per_binary = 1.0 + 0.8 * f_sn
    nb = int(np.ceil(n_orbits / exp_per_binary * 1.002)) + 16
    pop = make_population(nb, **kw)
    while pop.batch.n_orbits < n_orbits:
        nb = int(nb * 1.01) + 16
        pop = make_population(nb, **kw)
    if pop.batch.n_orbits > n_orbits:
        pop.batch = pop.batch.head(n_orbits)
    return pop


def population_tables(pop: SyntheticPopulation, bin_num0: int = 0):
    """The same synthetic population as the TABLES ``Population.perform_galactic_evolution`` reads
    (pop.py:1195-1313): ``(InitialGalaxy, bpp, kick_info, initial_binaries)``, pandas frames
    indexed by ``bin_num``, from which ``events.identify_events`` recovers exactly ``pop.primary_events`` /
    ``pop.secondary_events``.  The first supernova of a system is star 1's (``evol_type`` 15), the second star 2's
    (16); a disrupted system gets an ``evol_type`` 11 row and ends with ``sep < 0`` (pop.py:861-874).
    INPUT GENERATION ONLY (stands in for COSMIC)."""
    import pandas as pd
    from .pop import InitialGalaxy
    n = len(pop.tau_gyr)
    bins = np.arange(n, dtype=np.int64) + int(bin_num0)
    pe, se = pop.primary_events, pop.secondary_events
    K_ = len(pe)
    first = np.ones(K_, dtype=bool)
    first[1:] = pe.owner[1:] != pe.owner[:-1]
    star = np.where(first, 1, 2)
    last = np.ones(K_, dtype=bool)
    last[:-1] = pe.owner[1:] != pe.owner[:-1]
    row_disrupted = last & pop.disrupted[pe.owner]
    # secondary kick of every SN row: what the secondary table holds for disrupted systems, zero elsewhere (unused)
    dv2 = np.zeros((K_, 3))
    if len(se):
        key_p = pe.owner.astype(np.int64) * 4 + star
        first_s = np.ones(len(se), dtype=bool)
        first_s[1:] = se.owner[1:] != se.owner[:-1]
        key_s = se.owner.astype(np.int64) * 4 + np.where(first_s, 1, 2)
        pos = np.searchsorted(key_p, key_s)
        dv2[pos] = se.dv
    kick_info = pd.DataFrame({
        "star": star.astype(np.float64), "disrupted": row_disrupted.astype(np.float64),
        "delta_vsysx_1": pe.dv[:, 0], "delta_vsysy_1": pe.dv[:, 1], "delta_vsysz_1": pe.dv[:, 2],
        "delta_vsysx_2": dv2[:, 0], "delta_vsysy_2": dv2[:, 1], "delta_vsysz_2": dv2[:, 2],
        "bin_num": bins[pe.owner]})
    # COSMIC also writes star = 0 rows for systems without a supernova: one per system, ignored by the joins
    kick_info = pd.concat([pd.DataFrame({c: (bins if c == "bin_num" else np.zeros(n)) for c in kick_info.columns}),
                           kick_info], ignore_index=True)
    kick_info = kick_info.sort_values(["bin_num", "star"], kind="stable")
    kick_info.index = kick_info["bin_num"].values

    # bpp: birth row, the SN rows, the disruption row, the final row -- sorted by system, then time
    dis_rows = np.flatnonzero(row_disrupted)
    b_bin = np.concatenate([bins, bins[pe.owner], bins[pe.owner[dis_rows]], bins])
    b_t = np.concatenate([np.zeros(n), pe.tphys, pe.tphys[dis_rows], pop.tau_gyr * 1000.0])
    b_type = np.concatenate([np.ones(n), np.where(star == 1, 15.0, 16.0), np.full(len(dis_rows), 11.0), np.full(n, 10.0)])
    b_sep = np.concatenate([np.full(n, 100.0), np.where(row_disrupted, -1.0, 50.0), np.full(len(dis_rows), -1.0),
                            np.where(pop.disrupted, -1.0, 10.0)])
    b_ord = np.concatenate([np.zeros(n), np.ones(K_), np.full(len(dis_rows), 2.0), np.full(n, 3.0)])
    o = np.lexsort((b_ord, b_t, b_bin))
    bpp = pd.DataFrame({"tphys": b_t[o], "evol_type": b_type[o], "sep": b_sep[o], "bin_num": b_bin[o]})
    bpp.index = bpp["bin_num"].values

    ph = np.zeros((n, 2))
    ic = np.zeros((n, 2))
    ph[pe.owner, star - 1] = pe.phase
    ic[pe.owner, star - 1] = pe.inc
    initial_binaries = pd.DataFrame({"bin_num": bins, "metallicity": np.full(n, 0.02), "phase_sn_1": ph[:, 0],
                                     "phase_sn_2": ph[:, 1], "inc_sn_1": ic[:, 0], "inc_sn_2": ic[:, 1]})
    initial_binaries.index = bins
    galaxy = InitialGalaxy(pop.tau_gyr, pop.pos_kpc[0], pop.pos_kpc[1], pop.pos_kpc[2],
                           pop.vel_kms[0], pop.vel_kms[1], pop.vel_kms[2])
    return galaxy, bpp, kick_info, initial_binaries
