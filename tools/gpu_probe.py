"""First-contact probe on the GPU box: smoke, math self-test, FP64 peak, quick throughput."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits, make_population
import __graft_entry__ as g

g.smoke()
ctx = cb.Context(devices=[0])
print("math selftest", ctx.selftest_math())
for ms in (20, 100, 400):
    print("fp64 peak TFLOP/s", ms, ctx.measure_fp64_peak(0, ms))
n = int(os.environ.get("N_ORBITS", 1000000))
pop = population_with_n_orbits(n, cfg_id=4, horizon_gyr=1.0)
b = pop.batch
ctx.set_potential(pop.potential)
for integ, name in ((_lib.LEAPFROG, "leapfrog"),):
    for rep in range(3):
        t = time.time()
        res = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=integ)
        wall = time.time() - t
        steps = int(res["stats"][3])
        kms = ctx.last_kernel_ms(0)
        print(name, "orbits", b.n_orbits, "steps", steps, "kernel ms", kms, "total kernel ms", ctx.last_total_kernel_ms(0),
              "wall s", wall, "steps/s (kernel)", steps / (kms * 1e-3), "status!=0", int((res["status"] != 0).sum()))
sub = b.take(np.arange(min(n, 200000)))
for rep in range(2):
    res = ctx.integrate(sub.w0, sub.t1, sub.t2, sub.dt, sub.to_myr, sub.flags, sub.ev_off, sub.ev_time, sub.ev_dv, integrator=_lib.DOP853)
    kms = ctx.last_kernel_ms(0)
    print("dop853 orbits", sub.n_orbits, "stats", res["stats"], "kernel ms", kms, "intervals/s", res["stats"][3] / (kms * 1e-3),
          "grad evals/s", res["stats"][2] / (kms * 1e-3), "failed", int((res["status"] != 0).sum()))
