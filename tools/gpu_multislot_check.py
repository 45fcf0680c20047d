import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import cogsworth_b200 as cb
from cogsworth_b200 import _lib, kicks
from cogsworth_b200.synthetic import make_population, population_with_n_orbits
from cogsworth_b200.batch import build_orbit_batch
pop = population_with_n_orbits(1_500_000, cfg_id=4, horizon_gyr=0.2, f_sn=1.0)
one = cb.Context(devices=[0]); one.set_potential(pop.potential)
ref = one.integrate_batch(pop.batch)
for ns in (3, 8):
    many = cb.Context(devices=[0] * ns); many.set_potential(pop.potential)
    assert many.n_devices == ns
    for rep in range(2):
        r = many.integrate_batch(pop.batch)
        assert np.array_equal(r["w_final"], ref["w_final"]) and np.array_equal(r["status"], ref["status"]) and np.array_equal(r["ev_step"], ref["ev_step"]) and np.array_equal(r["n_nodes"], ref["n_nodes"])
    # sorted-by-age population, plain numpy (staged) arrays
    srt = make_population(200_000, cfg_id=52, horizon_gyr=(0.02, 0.4))
    o = np.argsort(srt.tau_gyr, kind="stable")
    sb = build_orbit_batch(srt.pos_kpc[:, o], srt.vel_kms[:, o], srt.tau_gyr[o], 12.0, 1.0, np.zeros(len(o), dtype=bool))
    many.set_potential(srt.potential); one.set_potential(srt.potential)
    a, b = one.integrate_batch(sb), many.integrate_batch(sb)
    assert np.array_equal(a["w_final"], b["w_final"]) and np.array_equal(a["n_nodes"], b["n_nodes"])
    ms = [many.last_kernel_ms(k) for k in range(ns)]
    print(ns, "slots on one device: same bits; kernel ms per slot", " ".join("%.2f" % x for x in ms), flush=True)
    one.set_potential(pop.potential)
    many.close()
print("ok")
