"""End-to-end time of ctx.integrate with plain numpy (pageable) arrays vs pinned tensors, 10^7 orbits."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pop = population_with_n_orbits(n, cfg_id=4, horizon_gyr=1.0)
b = pop.batch
ctx = cb.Context(devices=[0]); ctx.set_potential(pop.potential)
args = (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
for rep in range(3):
    t0 = time.perf_counter()
    r = ctx.integrate(*args, integrator=_lib.LEAPFROG)
    dt = time.perf_counter() - t0
    print("numpy (pageable) ctx.integrate: %.1f ms -> %.3e steps/s" % (dt * 1e3, 1000.0 * n / dt), flush=True)
