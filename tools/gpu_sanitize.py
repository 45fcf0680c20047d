"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): every kernel, every trajectory path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import make_population
ctx = cb.Context(devices=[0])
pop = make_population(700, cfg_id=77, horizon_gyr=None if "--ragged" in sys.argv else 0.12)
b = pop.batch
if "--ragged" in sys.argv:                       # cut the horizons so the run stays short under the sanitizer
    b.t1 = np.maximum(b.t1, b.t2 - 150.0 * np.where((b.flags & 1) == 1, 3.15576e13, 1.0))
ctx.set_potential(pop.potential)
args = (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
for integ in (_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE):
    for store in (False, True):
        r = ctx.integrate(*args, integrator=integ, store_all=store)
        print("integrator", integ, "store_all", store, "ok", int((r["status"] == 0).sum()), "of", b.n_orbits, flush=True)
q = np.random.default_rng(0).normal(0, 6, (3, 1000))
ctx.gradient(q); ctx.potential_value(q); ctx.circular_velocity(q)
ctx.close()
print("done")
