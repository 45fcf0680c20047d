"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel, every
trajectory path, both host data paths (page-locked + compact forms with the unpack kernel, staged plain arrays).

    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py [--ragged]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.potential import LM10Potential, evolving_milky_way
from cogsworth_b200.synthetic import make_population

ragged = "--ragged" in sys.argv
ctx = cb.Context(devices=[0])
pop = make_population(700, cfg_id=77, horizon_gyr=None if ragged else 0.12)
b = pop.batch
if ragged:                                       # cut the horizons so the run stays short under the sanitizer
    b.t1 = np.maximum(b.t1, b.t2 - 150.0 * np.where((b.flags & 1) == 1, 3.15576e13, 1.0))
    pop2 = make_population(500, cfg_id=78, horizon_gyr=(0.01, 0.14))     # a compact batch with ragged horizons
else:
    pop2 = pop
ctx.set_potential(pop.potential)
args = (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)
for integ in (_lib.LEAPFROG, _lib.DOP853, _lib.DOP853_DENSE):
    for store in (False, True):
        r = ctx.integrate(*args, integrator=integ, store_all=store)
        print("plain   integrator", integ, "store_all", store, "ok", int((r["status"] == 0).sum()), "of", b.n_orbits, flush=True)
        os.environ["CW_HOST_CHUNK"] = "300"      # several chunks per stream lane: arena and staging re-use
        r2 = ctx.integrate_batch(pop2.batch, integrator=integ, store_all=store)
        os.environ["CW_FORCE_STAGING"] = "1"
        r3 = ctx.integrate_batch(pop2.batch, integrator=integ, store_all=store)
        del os.environ["CW_FORCE_STAGING"], os.environ["CW_HOST_CHUNK"]
        assert np.array_equal(r2["w_final"], r3["w_final"])
        print("compact integrator", integ, "store_all", store, "ok", int((r2["status"] == 0).sum()), "of", pop2.batch.n_orbits, flush=True)
# the other kernel kinds: v1, generic, separable time dependence, general (rotated halo)
from cogsworth_b200.potential import CompositePotential, MilkyWayPotential
small = make_population(200, cfg_id=79, horizon_gyr=0.06).batch
for name, pot in (("v1", MilkyWayPotential("v1")),
                  ("generic", CompositePotential(names=["b", "d", "h"], types=[2, 1, 3], params=[(4e9, 0.8, 0.0), (5e10, 3.2, 0.25), (6e11, 17.0, 0.0)])),
                  ("evolving", evolving_milky_way(np.linspace(0.0, 12000.0, 5), np.linspace(0.1, 1.0, 5))),
                  ("lm10", LM10Potential())):
    ctx.set_potential(pot)
    for integ in (_lib.LEAPFROG, _lib.DOP853):
        r = ctx.integrate_batch(small, integrator=integ, store_all=(integ == _lib.LEAPFROG))
        print(name, "integrator", integ, "ok", int((r["status"] == 0).sum()), "of", small.n_orbits, flush=True)
ctx.set_potential(pop.potential)
q = np.random.default_rng(0).normal(0, 6, (3, 1000))
ctx.gradient(q); ctx.potential_value(q); ctx.circular_velocity(q)
ctx.count_nodes(None, b.t1, b.t2, b.dt, b.to_myr, b.flags)
ctx.close()
print("done")
