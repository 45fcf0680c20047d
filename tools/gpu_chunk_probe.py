"""Host-mode chunking experiment at a per-rank size of the 8-GPU run (1.25e6 orbits): CW_HOST_CHUNK sweep (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits

n_orbits = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_250_000
pop = population_with_n_orbits(n_orbits, cfg_id=4, horizon_gyr=1.0, f_sn=1.0)
b = pop.batch
ctx = cb.Context(devices=[0])
ctx.set_potential(pop.potential)
n = b.n_orbits
a = b.abi_arrays()
compact = dict(n_w0=a["n_w0"], w0_index=a["w0_index"], grid_class=a["grid_class"], classes=a["classes"], ev_off32=a["ev_off32"])
wf, st = _lib.pinned_empty((6, n)), _lib.pinned_empty(n, np.int32)


def run():
    ctx.integrate_raw(n, a["w0"], a["t1"], None, None, None, None, None, a["ev_time"], a["ev_dv"], wf, st, **compact)


def timeit(label):
    run(); run()
    ts = []
    for _ in range(7):
        t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
    print(f"{label:40s} best {min(ts) * 1e3:7.3f} ms  median {np.median(ts) * 1e3:7.3f} ms   kernels {ctx.last_kernel_ms(0):7.3f} ms  launches {ctx.last_launch_count}", flush=True)


Q = 151552
timeit("default")
for c in (Q // 2, Q, 2 * Q, 3 * Q, 4 * Q, 250000, 312500, 416667, 625000, n):
    os.environ["CW_HOST_CHUNK"] = str(c)
    timeit(f"CW_HOST_CHUNK={c} ({c / Q:.2f} waves)")
del os.environ["CW_HOST_CHUNK"]
os.environ["CW_NO_WAVE_CHUNKS"] = "1"
timeit("CW_NO_WAVE_CHUNKS")
del os.environ["CW_NO_WAVE_CHUNKS"]
# device-resident reference
import torch
dev = torch.device("cuda", 0)
to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
D = [to(x) for x in (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv)]
wfd, std = torch.empty((6, n), dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
def rund():
    ctx.integrate_raw(n, *D, wfd, std, mem_space=_lib.MEM_DEVICE)
rund(); rund()
ts = []
for _ in range(7):
    t = time.perf_counter(); rund(); ts.append(time.perf_counter() - t)
print(f"device-resident                          best {min(ts) * 1e3:7.3f} ms  kernels {ctx.last_kernel_ms(0):7.3f}")
os.environ["CW_DEBUG"] = "1"
run()
