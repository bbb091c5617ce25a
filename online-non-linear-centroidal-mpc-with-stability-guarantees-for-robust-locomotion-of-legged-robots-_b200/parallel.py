"""Multi-GPU: instances are independent, so the batch is cut into contiguous shards, one process per GPU, and no
collective touches the hot path.  The only exchange is a handful of scalars for statistics (converged count,
iterations, slowest rank's time) through torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` instances for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_arrays(arrays, rank: int, world: int):
    """Slice every instance-major array (leading dimension = instance) to this rank's shard."""
    lo, hi = shard_range(len(arrays[0]), rank, world)
    return [np.ascontiguousarray(a[lo:hi]) for a in arrays], (lo, hi)


def gather_stats(local: dict, dist=None, device=None) -> dict:
    """Whole-job statistics: sums of counters and the maximum of times.  Keys ending in `_ms` or `_s` are reduced with
    MAX (the job is as slow as its slowest rank), `*_max` with MAX, everything else with SUM."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    import torch
    keys = sorted(local)
    mx = [k for k in keys if k.endswith(("_ms", "_s", "_max"))]
    sm = [k for k in keys if k not in mx]
    out = {}
    for ks, op in ((mx, dist.ReduceOp.MAX), (sm, dist.ReduceOp.SUM)):
        if not ks:
            continue
        t = torch.tensor([float(local[k]) for k in ks], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        out.update({k: float(v) for k, v in zip(ks, t.cpu())})
    return out
