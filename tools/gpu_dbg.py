import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
ctx = cb.Context(devices=[0])
for f_sn, hz in ((0.0, 1.0), (0.0, 1.001), (0.0, 1.004), (1.0, 1.0)):
    pop = population_with_n_orbits(200000, cfg_id=3, horizon_gyr=hz, f_sn=f_sn)
    b = pop.batch
    ctx.set_potential(pop.potential)
    r = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG, store_all=True, with_t=False)
    st = [int(x) for x in r["stats"]]
    print("f_sn", f_sn, "hz", hz, "n_nodes", r["n_nodes"][:2], "parks", st[0], "generic events", st[1] // 1000000, "generic lanes", st[1] % 1000000,
          "lockstep", st[2] % 1000000000, "unpark_direct", st[2] // 1000000000, "steps", st[3], "kernel ms", ctx.last_kernel_ms(0), flush=True)
