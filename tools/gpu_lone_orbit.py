"""One warp of nucleus-bound orbits through the DOPRI853 kernel: what does a lone warp's step attempt cost, and on what
does it wait?  (GPU box; run it under `ncu --set full -k regex:dop853_kernel` for the stall breakdown.)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.potential import MilkyWayPotential

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
span = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
ctx = cb.Context(devices=[0])
pot = MilkyWayPotential("v2")
ctx.set_potential(pot)
r = 0.03                                                   # kpc: inside the nucleus, period ~ 1 Myr
vc = float(ctx.circular_velocity(np.array([[r], [0.0], [0.0]]))[0])
w0 = np.zeros((6, n))
w0[0] = r * (1.0 + 1e-3 * np.arange(n))
w0[4] = vc * 0.9
w0[5] = vc * 0.1
for rep in range(3):
    t0 = time.perf_counter()
    out = ctx.integrate(w0, np.zeros(n), np.full(n, span), np.ones(n), integrator=_lib.DOP853)
    dt = time.perf_counter() - t0
    att = ctx.last_max_attempts
    print(f"{n} orbits x {span:.0f} Myr: kernel {ctx.last_kernel_ms(0):8.2f} ms, wall {1e3 * dt:8.2f} ms, most attempts of one orbit "
          f"{att}, {1e3 * ctx.last_kernel_ms(0) / max(att, 1):.3f} us per attempt, status ok {(out['status'] == 0).all()}", flush=True)
