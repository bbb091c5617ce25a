"""N > 1 path on CPU: world_size-2 gloo processes shard a batch, "solve" their shard (a stand-in that marks every
instance with its global index), and reduce the statistics exactly as bench.py does on the GPUs."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    import cmpc_loader
    pkg = cmpc_loader.load()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 4097                                            # ragged on purpose
    x0 = np.arange(total * 20, dtype=float).reshape(total, 20)
    (mine,), (lo, hi) = pkg.shard_arrays([x0], rank, world)
    assert mine.shape[0] == hi - lo and mine[0, 0] == lo * 20
    stats = pkg.gather_stats({"converged": hi - lo, "iters": 10 * (hi - lo), "step_ms": 5.0 + rank, "viol_max": 1e-9 * (rank + 1)}, dist)
    q.put((rank, lo, hi, stats))
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert [p.exitcode for p in ps] == [0, 0]
    (r0, lo0, hi0, s0), (r1, lo1, hi1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 2049, 2049, 4097)
    for s in (s0, s1):
        assert s["converged"] == 4097 and s["iters"] == 40970 and s["step_ms"] == 6.0 and s["viol_max"] == 2e-9


@pytest.mark.parametrize("total,world", [(4096, 8), (65536, 8), (5, 8), (0, 2), (16384, 4)])
def test_shard_ranges_partition_the_batch(pkg, total, world):
    r = [pkg.shard_range(total, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == total
    assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1
