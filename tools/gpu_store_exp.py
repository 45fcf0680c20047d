"""store_all kernel time vs orbit length / kick presence (GPU box): is the kicked slowdown the kicks
or the trajectory stride / column phase pattern?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
ctx = cb.Context(devices=[0])
N = 300000
for f_sn, hz in ((0.0, 1.0), (0.0, 1.001), (0.0, 1.002), (0.0, 1.004), (0.0, 1.008), (0.0, 1.016), (1.0, 1.0), (1.0, 1.015)):
    pop = population_with_n_orbits(N, cfg_id=3, horizon_gyr=hz, f_sn=f_sn)
    b = pop.batch
    ctx.set_potential(pop.potential)
    ms = []
    for rep in range(3):
        r = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG,
                          store_all=True, with_t=False)
        ms.append(ctx.last_kernel_ms(0))
    steps = int((r["n_nodes"].astype(np.int64)).sum())
    print(f"f_sn {f_sn} horizon {hz} n_nodes {r['n_nodes'][:2]} events {b.n_events} kernel ms {min(ms):.3f} "
          f"-> {48.0 * steps / (min(ms) * 1e-3) / 1e9:.0f} GB/s", flush=True)
