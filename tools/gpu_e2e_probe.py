"""e2e (host-buffer) vs device-resident time at the per-rank size of an 8-GPU run, with the shard timeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import population_with_n_orbits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
pop = population_with_n_orbits(n, cfg_id=4, horizon_gyr=1.0)
b = pop.batch
ctx = cb.Context(devices=[0]); ctx.set_potential(pop.potential)
pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).pin_memory()
f64, i32, i64 = torch.float64, torch.int32, torch.int64
H = dict(w0=pin(b.w0, f64), t1=pin(b.t1, f64), t2=pin(b.t2, f64), dt=pin(b.dt, f64), to_myr=pin(b.to_myr, f64), flags=pin(b.flags, i32),
         ev_off=pin(b.ev_off, i64), ev_time=pin(b.ev_time, f64), ev_dv=pin(b.ev_dv, f64))
wf = torch.empty((6, n), dtype=f64).pin_memory(); st = torch.empty(n, dtype=i32).pin_memory(); stats = np.zeros(4, dtype=np.int64)
def step():
    ctx.integrate_raw(n, H["w0"], H["t1"], H["t2"], H["dt"], H["to_myr"], H["flags"], H["ev_off"], H["ev_time"], H["ev_dv"], wf, st, None, None,
                      None, None, None, None, stats, integrator=_lib.LEAPFROG, mem_space=_lib.MEM_HOST)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
dt = (time.perf_counter() - t0) / 5
print("e2e ms/step %.3f  (kernel sum %.3f ms)" % (dt * 1e3, ctx.last_kernel_ms(0)), flush=True)
os.environ["CW_DEBUG"] = "1"
step()
