"""Gene-pair sharding across the GPUs of one box.

One process per GPU (torchrun).  Jobs are independent, so each rank evaluates one contiguous slab of the
job list on its own device and a single all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests)
rebuilds the full result vector on every rank -- the only collective on the path.  Replaces the
reference's unusable multiprocessing.Pool branch (causal_network.py:181,:286).
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the initialised torch.distributed group, else (0, 1)."""
    try:
        import torch.distributed as dist
    except Exception:      # torch absent: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slab(n_jobs, rank, world_size):
    """Contiguous [lo, hi) of `n_jobs` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_jobs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_evaluate(n_jobs, evaluate, device=None):
    """Evaluate jobs [0, n_jobs) across the process group.

    evaluate(lo, hi) -> float64 array of hi-lo results for this rank's slab (runs on this rank's GPU).
    Returns the full (n_jobs,) float64 array on every rank.  Single process: evaluate(0, n_jobs).
    """
    rank, ws = world()
    if ws == 1:
        return np.asarray(evaluate(0, n_jobs), dtype=np.float64)
    import torch
    import torch.distributed as dist
    lo, hi = slab(n_jobs, rank, ws)
    local = np.asarray(evaluate(lo, hi), dtype=np.float64)
    assert local.shape == (hi - lo,)
    width = slab(n_jobs, 0, ws)[1]                       # largest slab; pad so all_gather_into_tensor applies
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(width, dtype=torch.float64, device=device)
    buf[:hi - lo] = torch.from_numpy(local).to(device)
    gathered = torch.empty(width * ws, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(gathered, buf)
    g = gathered.cpu().numpy().reshape(ws, width)
    out = np.empty(n_jobs, dtype=np.float64)
    for r in range(ws):
        a, b = slab(n_jobs, r, ws)
        out[a:b] = g[r, :b - a]
    return out
