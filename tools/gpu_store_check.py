"""Strict store_all check: every stored sample against the oracle, column by column (not quantiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb, oracle as O
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import make_population
ctx = cb.Context(devices=[0])
for label, pop in (("ragged (Wagg horizons)", make_population(300, cfg_id=21)),
                   ("mixed 250/251 nodes", make_population(600, cfg_id=3, horizon_gyr=0.25)),
                   ("kicked 1001", make_population(400, cfg_id=4, horizon_gyr=1.0))):
    b = pop.batch
    ctx.set_potential(pop.potential)
    for integ in (_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE):
        g = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ, store_all=True)
        o = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                              integrator=integ, traj_off=g["traj_off"])
        off = g["traj_off"]
        # per-orbit: a smooth (non-chaotic) orbit agrees everywhere; a corrupted column shows as an isolated spike
        d = np.abs(g["traj_pos"] - o["traj_pos"]).max(axis=0) + 1e3 * np.abs(g["traj_vel"] - o["traj_vel"]).max(axis=0)
        bad_orbits = 0; spikes = 0
        for i in range(b.n_orbits):
            di = d[off[i]:off[i + 1]]
            if len(di) < 3: continue
            # spike = sample much worse than both neighbours and than the orbit's final error
            inner = di[1:-1]
            sp = (inner > 1e-6) & (inner > 100 * np.maximum(di[:-2], di[2:]))
            spikes += int(sp.sum())
            if di.max() > 1e-6: bad_orbits += 1
        print(f"{label:24s} integ {integ}: orbits {b.n_orbits}, columns {off[-1]}, max diff {d.max():.2e}, orbits>1e-6: {bad_orbits}, isolated spikes: {spikes}", flush=True)
