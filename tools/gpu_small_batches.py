"""Leapfrog kernel time for small and mid-size batches (device time of the integrator kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
ctx = cb.Context(devices=[0])
for n in (2000, 10000, 20000, 50000, 100000, 140000, 300000):
    pop = population_with_n_orbits(n, cfg_id=4, horizon_gyr=1.0)
    b = pop.batch
    ctx.set_potential(pop.potential)
    ms = []
    for _ in range(4):
        ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG)
        ms.append(ctx.last_kernel_ms(0))
    print(f"n {n:7d}: kernel {min(ms):7.3f} ms  -> {1000.0 * n / (min(ms) * 1e-3):.3e} steps/s", flush=True)
