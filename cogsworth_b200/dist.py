"""Multi-GPU plumbing: one process per GPU, orbits sharded by index, NO collective on the data path.

The reference parallelises this stage with ``multiprocessing.Pool.starmap`` over independent orbits
(``src/cogsworth/pop.py:1245-1250``); here the same independence means each rank integrates its own
contiguous slice and the only communication is the timing barrier / reduction of the benchmark and an
optional gather of the 48-byte final states.  Backend-agnostic (``nccl`` on GPUs, ``gloo`` in CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """[a, b) of rank ``rank``: the same balanced contiguous split cw_integrate uses across device slots."""
    return n * rank // world, n * (rank + 1) // world


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def reduce_scalar(x: float, op: str = "sum", device=None) -> float:
    """all-reduce of one float ('sum' | 'max'); identity without a process group."""
    d = _dist()
    if d is None:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.SUM if op == "sum" else d.ReduceOp.MAX)
    return float(t.item())


def gather_final_states(w_final_local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Optional host-side gather of every rank's (6, n_local) final states into (6, n_total) on all ranks
    ('only a final host-side gather').  Shards must follow shard_range order."""
    d = _dist()
    if d is None:
        return w_final_local
    import torch
    rank, ws = d.get_rank(), d.get_world_size()
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    nmax = max(b - a for a, b in sizes)
    buf = torch.zeros((6, nmax), dtype=torch.float64, device=device)
    buf[:, :w_final_local.shape[1]] = torch.from_numpy(np.ascontiguousarray(w_final_local)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(ws)]
    d.all_gather(out, buf)
    full = np.empty((6, n_total))
    for (a, b), t in zip(sizes, out):
        full[:, a:b] = t[:, :b - a].cpu().numpy()
    return full
