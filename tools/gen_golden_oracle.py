#!/usr/bin/env python
"""Freeze oracle outputs on small seeded cases into tests/golden/oracle_cases.npz.

These are NOT reference (gala) outputs -- gala cannot run here -- they pin the oracle itself (and
through it the CUDA path) against drift: tests/test_oracle_golden.py re-runs the oracle against
them on the CPU, tests/test_gpu_parity.py compares the CUDA path with them on the B200.
Cases: config 1 (MW v1, leapfrog, 200 x 1 Myr, kicks) and DOPRI853 / MW v2 variants in both forms (restart per
interval, dense output), 96 binaries.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from cogsworth_b200.potential import MilkyWayPotential  # noqa: E402
from cogsworth_b200.synthetic import make_population  # noqa: E402

out = {}
for name, ver, integ in (("cfg1_leapfrog_v1", "v1", O.LEAPFROG), ("cfg1_dop853_v2", "v2", O.DOP853),
                         ("cfg1_dop853dense_v2", "v2", O.DOP853_DENSE)):
    pot = MilkyWayPotential(ver)
    pop = make_population(96, cfg_id=1, potential=pot, horizon_gyr=0.2)
    b = pop.batch
    r = O.integrate_batch(pot, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                          integrator=integ, n_threads=1)
    assert (r["status"] == 0).all()
    out[name + "_w0"] = b.w0
    out[name + "_w_final"] = r["w_final"]
    out[name + "_ev_step"] = r["ev_step"]
    out[name + "_n_nodes"] = r["n_nodes"]
    out[name + "_ev_time"] = b.ev_time
# the general potentials (rotation, per-component time interpolation): birth velocities drawn in the
# static Milky Way, orbits integrated in the named potential
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _cases import general_cases  # noqa: E402

for name, pot, integ, horizon in general_cases():
    pop = make_population(48, cfg_id=1, potential=MilkyWayPotential("v2"), horizon_gyr=horizon)
    b = pop.batch
    r = O.integrate_batch(pot, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                          integrator=integ, n_threads=1)
    assert (r["status"] == 0).all()
    out[name + "_w_final"] = r["w_final"]
    out[name + "_ev_step"] = r["ev_step"]
    out[name + "_n_nodes"] = r["n_nodes"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_cases.npz"), **out)
print({k: v.shape for k, v in out.items()})
