"""FP64-side rate of the final-state leapfrog kernel at reduced occupancy (CW_LF_MAX_CTAS): how many orbit-steps/s can
8 / 16 / 24 / 32 warps per SM sustain?  Sizes the lanes-per-SM budget of the trajectory kernel (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
pop = population_with_n_orbits(2_000_000, cfg_id=4, horizon_gyr=1.0, f_sn=0.0)
b = pop.batch
ctx = cb.Context(devices=[0]); ctx.set_potential(pop.potential)
dev = torch.device("cuda", 0)
to = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(dev)
n = b.n_orbits
D = [to(x) for x in (b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags)] + [None, None, None]
wf, st = torch.empty((6, n), dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
for ctas in (4, 3, 2, 1):
    os.environ["CW_LF_MAX_CTAS"] = str(ctas)
    for _ in range(2):
        ctx.integrate_raw(n, *D, wf, st, mem_space=_lib.MEM_DEVICE)
    ms = ctx.last_kernel_ms(0)
    print(f"{ctas} CTAs x 256 threads per SM ({8 * ctas} warps): kernel {ms:.3f} ms  {n * 999 / ms / 1e6:.4g} G orbit-steps/s", flush=True)
