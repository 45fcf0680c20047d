"""End-to-end timings of the CW_MEM_HOST path variants at config-4 size (GPU box; not a bench line).

usage: python tools/gpu_hostpath_probe.py [n_orbits] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import _lib, kicks
from cogsworth_b200.synthetic import population_with_n_orbits

n_orbits = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.perf_counter()
pop = population_with_n_orbits(n_orbits, cfg_id=4, horizon_gyr=1.0, f_sn=1.0)
b = pop.batch
print(f"population: {b.n_orbits} orbits, {b.n_events} events, compact={b.compact is not None}, "
      f"built in {time.perf_counter() - t0:.2f} s", flush=True)
ctx = cb.Context(devices=[0])
ctx.set_potential(pop.potential)
n, K = b.n_orbits, b.n_events
steps = 1000 * n


def timeit(name, fn, nrep=reps, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(nrep):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    best, med = min(ts), float(np.median(ts))
    print(f"{name:58s} best {best * 1e3:8.2f} ms  median {med * 1e3:8.2f} ms  {steps / best:.3e} steps/s  "
          f"kernel {ctx.last_kernel_ms(0):7.2f} ms", flush=True)
    return best


a = b.abi_arrays()
compact = dict(n_w0=a["n_w0"], w0_index=a["w0_index"], grid_class=a["grid_class"], classes=a["classes"],
               ev_off32=a["ev_off32"])
wf, st = _lib.pinned_empty((6, n)), _lib.pinned_empty(n, np.int32)
h2d_compact = sum(x.nbytes for x in (a["w0"], a["t1"], a["w0_index"], a["grid_class"], a["ev_off32"], a["ev_time"], a["ev_dv"]))
print(f"H2D bytes/orbit: compact {h2d_compact / n:.1f}", flush=True)
timeit("C-ABI, page-locked arrays, compact forms",
       lambda: ctx.integrate_raw(n, a["w0"], a["t1"], None, None, None, None, None, a["ev_time"], a["ev_dv"], wf, st,
                                 **compact))
ref = wf.copy()

# plain layout in page-locked memory (the round-1 e2e)
P = {}
for k in ("w0", "t1", "t2", "dt", "to_myr", "flags", "ev_off"):
    src = getattr(b, k)
    P[k] = _lib.pinned_empty(src.shape, src.dtype)
    P[k][...] = src
h2d_plain = sum(x.nbytes for x in P.values()) + b.ev_time.nbytes + b.ev_dv.nbytes
print(f"H2D bytes/orbit: plain {h2d_plain / n:.1f}", flush=True)
timeit("C-ABI, page-locked arrays, plain layout",
       lambda: ctx.integrate_raw(n, P["w0"], P["t1"], P["t2"], P["dt"], P["to_myr"], P["flags"], P["ev_off"], b.ev_time,
                                 b.ev_dv, wf, st))
assert np.array_equal(ref, wf)

# pageable numpy copies of everything
Q = {k: np.array(v, copy=True) for k, v in P.items()}
evt, evd = np.array(b.ev_time), np.array(b.ev_dv)
wf2, st2 = np.empty((6, n)), np.empty(n, dtype=np.int32)
timeit("C-ABI, pageable arrays (staged), plain layout",
       lambda: ctx.integrate_raw(n, Q["w0"], Q["t1"], Q["t2"], Q["dt"], Q["to_myr"], Q["flags"], Q["ev_off"], evt, evd,
                                 wf2, st2))
assert np.array_equal(ref, wf2)
timeit("C-ABI, pageable inputs, page-locked outputs, plain",
       lambda: ctx.integrate_raw(n, Q["w0"], Q["t1"], Q["t2"], Q["dt"], Q["to_myr"], Q["flags"], Q["ev_off"], evt, evd,
                                 wf, st))
ca = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in compact.items()}
w0c, t1c = np.array(a["w0"]), np.array(a["t1"])
timeit("C-ABI, pageable inputs (staged), compact, pinned outputs",
       lambda: ctx.integrate_raw(n, w0c, t1c, None, None, None, None, None, evt, evd, wf, st, **ca))
assert np.array_equal(ref, wf)

timeit("Context.integrate_batch(batch)  [pinned in + pooled out]", lambda: ctx.integrate_batch(b))
timeit("Context.integrate(plain numpy arrays)  [staged in]",
       lambda: ctx.integrate(Q["w0"], Q["t1"], Q["t2"], Q["dt"], Q["to_myr"], Q["flags"], Q["ev_off"], evt, evd))
timeit("kicks.integrate_orbits(batch, store_all=False)",
       lambda: kicks.integrate_orbits(b, potential=pop.potential, store_all=False, integrator="leapfrog", context=ctx))
from cogsworth_b200.batch import OrbitBatch
bq = OrbitBatch(Q["w0"], Q["t1"], Q["t2"], Q["dt"], Q["to_myr"], Q["flags"], Q["ev_off"], evt, evd)
timeit("kicks.integrate_orbits(plain numpy batch, store_all=False)",
       lambda: kicks.integrate_orbits(bq, potential=pop.potential, store_all=False, integrator="leapfrog", context=ctx))
os.environ["CW_DEBUG"] = "1"
ctx.integrate_batch(b)
ctx.close()
