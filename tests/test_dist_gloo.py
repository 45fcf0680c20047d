"""world_size = 2 on CPU (gloo): the N > 1 plumbing of bench.py / cogsworth_b200.dist -- index sharding
with no data-path collective, sum / max reductions, final gather.  The integration itself is done by the
CPU oracle here (no GPU in this container); on the B200 box the same code paths run with nccl."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import bench
    from cogsworth_b200 import dist as cd
    from cogsworth_b200.synthetic import make_population
    # one population known to every rank, sharded by index
    pop = make_population(60, cfg_id=8, horizon_gyr=0.05)
    b = pop.batch
    a, e = cd.shard_range(b.n_orbits, rank, world)
    sub = b.take(np.arange(a, e))
    r = O.integrate_batch(pop.potential, sub.w0, sub.t1, sub.t2, sub.dt, sub.to_myr, sub.flags, sub.ev_off,
                          sub.ev_time, sub.ev_dv, integrator=O.LEAPFROG, n_threads=1)
    steps_local = float((r["n_nodes"] - 1).sum())
    total_steps = cd.reduce_scalar(steps_local, "sum")
    slowest = cd.reduce_scalar(1.0 + rank, "max")
    full = cd.gather_final_states(r["w_final"], b.n_orbits)
    # the bench's own strong-scaling shards: rank-specific seeds, sizes add up to the workload
    wl = dict(bench.WORKLOADS["config4"])
    shard = bench.build_shard(wl, rank, world, "strong", 2e-5)
    n_total = cd.reduce_scalar(float(shard.batch.n_orbits), "sum")
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), full=full, total_steps=total_steps, slowest=slowest,
             n_total=n_total, first=shard.batch.w0[:, 0], rng=(a, e))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path, built):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import oracle as O
    from cogsworth_b200.synthetic import make_population
    pop = make_population(60, cfg_id=8, horizon_gyr=0.05)
    b = pop.batch
    ref = O.integrate_batch(pop.potential, b.w0, b.t1, b.t2, b.dt, b.to_myr, b.flags, b.ev_off, b.ev_time, b.ev_dv,
                            integrator=O.LEAPFROG, n_threads=1)
    r0, r1 = (np.load(tmp_path / f"rank{k}.npz") for k in range(world))
    # disjoint, covering shards; gathered result identical on both ranks and equal to the single-process run
    assert tuple(r0["rng"]) == (0, b.n_orbits // 2) and tuple(r1["rng"]) == (b.n_orbits // 2, b.n_orbits)
    assert np.array_equal(r0["full"], r1["full"]) and np.array_equal(r0["full"], ref["w_final"])
    assert r0["total_steps"] == r1["total_steps"] == float((ref["n_nodes"] - 1).sum())
    assert r0["slowest"] == r1["slowest"] == 2.0            # max over ranks
    # strong scaling: every rank builds its own seeded share of the 10^7-orbit workload
    assert r0["n_total"] == r1["n_total"] == 200.0
    assert not np.array_equal(r0["first"], r1["first"])
