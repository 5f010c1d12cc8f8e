"""Gene-pair sharding across the GPUs of one box.

One process per GPU (torchrun).  Jobs are independent, so each rank evaluates one contiguous slab of the
job list on its own device and a single all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests)
rebuilds the full result vector on every rank -- the only collective on the path.  Replaces the
reference's unusable multiprocessing.Pool branch (causal_network.py:181,:286).

On GPUs the slab's results never visit the host before the collective: the library writes them into the
all-gather's send buffer (`scribe_*_jobs_dev`), NCCL gathers device to device, and the full vector is read
back once.  A rank whose evaluation raises still takes part in the collective and reports a status word in
it, so every rank raises instead of hanging in the all-gather.
"""
from __future__ import annotations

import numpy as np


class ShardError(RuntimeError):
    """Another rank of the process group failed while evaluating its slab."""


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


def sharded_evaluate(n_jobs, evaluate, handle=None):
    """Evaluate jobs [0, n_jobs) across the process group.

    evaluate(lo, hi, out_dev_ptr=None): results of this rank's slab, computed on this rank's GPU.  With
    out_dev_ptr (an address in the memory of `handle`'s device, room for hi-lo float64) it writes them there and
    returns None; without it, it returns a float64 array of hi-lo values.
    Returns the full (n_jobs,) float64 array on every rank.  Single process: evaluate(0, n_jobs).
    """
    rank, ws = world()
    if ws == 1:
        return np.asarray(evaluate(0, n_jobs), dtype=np.float64)
    import torch
    import torch.distributed as dist
    lo, hi = slab(n_jobs, rank, ws)
    width = slab(n_jobs, 0, ws)[1] + 1                   # largest slab + one status word; equal sizes for all_gather_into_tensor
    on_gpu = dist.get_backend() == "nccl"
    if on_gpu:
        # the collective must run on the GPU the handle computes on (LOCAL_RANK by default), whatever the caller's
        # current device is: a buffer on another device would make two ranks share one GPU inside NCCL
        dev_index = getattr(handle, "device", None)
        if dev_index is None:
            dev_index = torch.cuda.current_device()
        device = torch.device("cuda", int(dev_index))
    else:
        device = torch.device("cpu")
    buf = torch.zeros(width, dtype=torch.float64, device=device)
    err = None
    try:
        if on_gpu and handle is not None and hasattr(handle, "device"):
            evaluate(lo, hi, buf.data_ptr())             # the library synchronises its stream before it returns
        else:
            local = np.asarray(evaluate(lo, hi), dtype=np.float64)
            assert local.shape == (hi - lo,)
            buf[:hi - lo] = torch.from_numpy(local).to(device)
    except BaseException as e:                           # noqa: BLE001 -- re-raised below, after the collective
        err = e
        buf.zero_()
        buf[width - 1] = 1.0
    gathered = torch.empty(width * ws, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(gathered, buf)
    g = gathered.cpu().numpy().reshape(ws, width)
    if err is not None:
        raise err
    failed = [r for r in range(ws) if g[r, width - 1] != 0.0]
    if failed:
        raise ShardError("rank(s) %s failed while evaluating their slab of the job table; see their traceback" % failed)
    out = np.empty(n_jobs, dtype=np.float64)
    for r in range(ws):
        a, b = slab(n_jobs, r, ws)
        out[a:b] = g[r, :b - a]
    return out
