import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
ctx = cb.Context(devices=[0])
for f_sn in (0.0, 1.0):
    pop = population_with_n_orbits(200000, cfg_id=3, horizon_gyr=1.0, f_sn=f_sn)
    b = pop.batch
    ctx.set_potential(pop.potential)
    r = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG, store_all=True, with_t=False)
    print("f_sn", f_sn, "n_nodes", r["n_nodes"][:3], "stats parks/generic/lockstep/steps", r["stats"], "kernel ms", ctx.last_kernel_ms(0))
