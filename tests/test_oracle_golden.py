"""The oracle and the synthetic generator reproduce the frozen cases of tests/golden/oracle_cases.npz
(made by tools/gen_golden_oracle.py) bit for bit -- guards both against silent drift."""
import os

import numpy as np
import pytest

from cogsworth_b200.potential import MilkyWayPotential
from cogsworth_b200.synthetic import make_population

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_cases.npz"))


@pytest.mark.parametrize("name,ver,integ", [("cfg1_leapfrog_v1", "v1", 0), ("cfg1_dop853_v2", "v2", 1),
                                           ("cfg1_dop853dense_v2", "v2", 2)])
def test_oracle_reproduces_golden(oracle, name, ver, integ):
    pot = MilkyWayPotential(ver)
    b = make_population(96, cfg_id=1, potential=pot, horizon_gyr=0.2).batch
    assert np.array_equal(b.w0, G[name + "_w0"])                 # seeded inputs are stable
    assert np.array_equal(b.ev_time, G[name + "_ev_time"])
    r = oracle.integrate_batch(pot, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                               integrator=integ)
    assert np.array_equal(r["ev_step"], G[name + "_ev_step"])
    assert np.array_equal(r["n_nodes"], G[name + "_n_nodes"])
    assert np.array_equal(r["w_final"], G[name + "_w_final"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_oracle_reproduces_golden_general_potentials(oracle, k):
    """LM10 (rotated halo), BovyMWPotential2014 (incomplete gamma function) and a per-component evolving Milky Way."""
    from _cases import general_cases
    name, pot, integ, horizon = general_cases()[k]
    b = make_population(48, cfg_id=1, potential=MilkyWayPotential("v2"), horizon_gyr=horizon).batch
    r = oracle.integrate_batch(pot, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                               integrator=integ)
    assert np.array_equal(r["ev_step"], G[name + "_ev_step"])
    assert np.array_equal(r["n_nodes"], G[name + "_n_nodes"])
    assert np.array_equal(r["w_final"], G[name + "_w_final"])
