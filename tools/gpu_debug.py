import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib
from cogsworth_b200.synthetic import make_population
import oracle as O
np.set_printoptions(linewidth=200)
pop = make_population(2048, cfg_id=99, horizon_gyr=0.2)
b = pop.batch
ctx = cb.Context(devices=[0])
print("selftest", ctx.selftest_math())
for ver in ("v2", "v1"):
    pot = cb.MilkyWayPotential(ver)
    ctx.set_potential(pot)
    q = np.ascontiguousarray(b.w0[:3, :512])
    g = ctx.gradient(q)
    go = np.array([O.gradient(pot, q[:, i]) for i in range(q.shape[1])]).T
    print(ver, "gradient max rel diff", (np.abs(g - go) / np.abs(go)).max())
    v = ctx.potential_value(q); vo = np.array([O.value(pot, q[:, i]) for i in range(q.shape[1])])
    print(ver, "value max rel", (np.abs(v - vo) / np.abs(vo)).max())
    vc = ctx.circular_velocity(q); vco = np.array([O.circular_velocity(pot, q[:, i]) for i in range(q.shape[1])])
    print(ver, "vcirc max rel", (np.abs(vc - vco) / np.abs(vco)).max())
ctx.set_potential(pop.potential)
# no-event orbits only, few steps
for nsteps in (1, 2, 10, 200):
    sub = b.take(np.arange(64))
    t1 = np.zeros(64); t2 = np.full(64, float(nsteps) + 0.5); dt = np.ones(64)
    res = ctx.integrate(sub.w0, t1, t2, dt, 1.0, 0, integrator=_lib.LEAPFROG)
    ref = O.integrate_batch(pop.potential, sub.w0, t1, t2, dt, 1.0, 0, integrator=O.LEAPFROG)
    rel = np.abs(res["w_final"] - ref["w_final"]) / np.maximum(np.abs(ref["w_final"]), 1e-3)
    print("noev nsteps", nsteps, "n_nodes", res["n_nodes"][:3], ref["n_nodes"][:3], "max rel", rel.max(), "median", np.median(rel))
res = ctx.integrate(b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=_lib.LEAPFROG)
ref = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv, integrator=O.LEAPFROG)
rel = (np.abs(res["w_final"] - ref["w_final"]) / np.maximum(np.abs(ref["w_final"]), 1e-3)).max(axis=0)
print("full: median", np.median(rel), "quantiles", np.quantile(rel, [0.5, 0.9, 0.99, 1.0]))
bad = np.argsort(rel)[-5:]
for i in bad:
    print(i, rel[i], "nev", b.ev_off[i + 1] - b.ev_off[i], "steps", res["ev_step"][b.ev_off[i]:b.ev_off[i + 1]], res["w_final"][:, i], ref["w_final"][:, i])
# trajectories for worst
i = int(bad[-1])
s1 = b.take([i])
r = ctx.integrate(s1.w0, s1.t1, s1.t2, s1.dt, s1.to_myr, s1.flags, s1.ev_off, s1.ev_time, s1.ev_dv, integrator=_lib.LEAPFROG, store_all=True)
o = O.integrate_batch(pop.potential, s1.w0, s1.t1, s1.t2, s1.dt, s1.to_myr, s1.flags, s1.ev_off, s1.ev_time, s1.ev_dv, integrator=O.LEAPFROG, traj_off=r["traj_off"])
d = np.abs(r["traj_pos"] - o["traj_pos"]).max(axis=0)
first = np.flatnonzero(d > 1e-9)
print("traj first divergence at", first[:5], "ev_step", r["ev_step"], o["ev_step"], "dt t", np.abs(r["traj_t"] - o["traj_t"]).max())
k = first[0] if len(first) else 0
print(r["traj_pos"][:, max(k - 2, 0):k + 3].T, o["traj_pos"][:, max(k - 2, 0):k + 3].T)
print(r["traj_vel"][:, max(k - 2, 0):k + 3].T, o["traj_vel"][:, max(k - 2, 0):k + 3].T)
