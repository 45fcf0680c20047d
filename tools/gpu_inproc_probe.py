"""One process, several GPUs: where does the wall clock of kicks.integrate_orbits go? (GPU box, >= 2 GPUs; not a bench line)

usage: python tools/gpu_inproc_probe.py [n_orbits] [reps] ["0;0,1;0,1,2,3"]
Runs the config-4 population through a context on device 0, on device 1, and on both, through the Python API and
through the bare C-ABI call, and prints the library's host-side milestones (CW_DEBUG) of one call of each.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cogsworth_b200 as cb
from cogsworth_b200 import _lib, kicks
from cogsworth_b200.synthetic import population_with_n_orbits

n_orbits = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ndev = torch.cuda.device_count()
pop = population_with_n_orbits(n_orbits, cfg_id=4, horizon_gyr=1.0, f_sn=1.0)
b = pop.batch
n = b.n_orbits
steps = 1000 * n
a = b.abi_arrays()
compact = dict(n_w0=a["n_w0"], w0_index=a["w0_index"], grid_class=a["grid_class"], classes=a["classes"],
               ev_off32=a["ev_off32"])
wf, st = _lib.pinned_empty((6, n)), _lib.pinned_empty(n, np.int32)
print(f"{n} orbits, {b.n_events} events, {ndev} device(s) visible", flush=True)


def timeit(name, fn, ctx, devs):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    km = " ".join(f"{ctx.last_kernel_ms(k):.2f}" for k in range(len(devs)))
    print(f"{name:40s} devices {devs}: best {min(ts) * 1e3:8.2f} ms  median {np.median(ts) * 1e3:8.2f} ms  "
          f"{steps / min(ts):.3e} steps/s  kernel ms per slot: {km}", flush=True)


if len(sys.argv) > 3:      # device sets, e.g. "0;0,1,2,3;0,1,2,3,4,5,6,7"
    sets = [[int(x) for x in grp.split(",")] for grp in sys.argv[3].split(";")]
else:
    sets = [[0]] + ([[1], [0, 1]] if ndev > 1 else []) + ([list(range(ndev))] if ndev > 2 else [])
for devs in sets:
    with cb.Context(devices=devs) as ctx:
        ctx.set_potential(pop.potential)
        timeit("C-ABI (cw_integrate, page-locked)", lambda: ctx.integrate_raw(
            n, a["w0"], a["t1"], None, None, None, None, None, a["ev_time"], a["ev_dv"], wf, st, **compact), ctx, devs)
        timeit("kicks.integrate_orbits", lambda: kicks.integrate_orbits(
            b, potential=pop.potential, store_all=False, integrator="leapfrog", context=ctx), ctx, devs)
        os.environ["CW_DEBUG"] = "1"
        sys.stderr.write(f"---- one traced C-ABI call on devices {devs}\n")
        sys.stderr.flush()
        ctx.integrate_raw(n, a["w0"], a["t1"], None, None, None, None, None, a["ev_time"], a["ev_dv"], wf, st, **compact)
        del os.environ["CW_DEBUG"]
