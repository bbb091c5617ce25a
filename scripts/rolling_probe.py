"""Rolling replay diagnostics: per-tick kernel time, iterations, tail."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cmpc_loader
pkg = cmpc_loader.load()
N, B, R = 20, 4096, 8
WM = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ticks, mass, k1, idx = bench.replay_workload(N, B, seed=0, back=2, ahead=R)
s = pkg.BatchSolver(N, B, device=0)
s.solve_host(*ticks[0], mass, k1, 0)
for r in range(1, R + 3):
    o = s.solve_host(*ticks[r], mass, k1, WM)
    st = s.last_stats()
    it = o["iters"]
    print("tick offset %+d: kernel %.1f ms  iters mean %.2f p99 %d max %d  nfact %.2f  conv %.4f  work-ideal %.1f ms" % (
        r - 2, st["kernel_ms"], it.mean(), np.percentile(it, 99), it.max(), st["nfact"] / B, (o["status"] == 0).mean(),
        st["kernel_ms"] * 0 + st["nfact"] * 1.31e6 / 444 / 1.965e6), flush=True)
# which instances are the stragglers?
s2 = pkg.BatchSolver(N, B, device=0)
s2.solve_host(*ticks[0], mass, k1, 0)
hist = {}
for r in range(1, R + 3):
    o = s2.solve_host(*ticks[r], mass, k1, WM)
    big = np.flatnonzero(o["iters"] > 80)
    for b in big:
        hist.setdefault(int(idx[b]) + r - 2, []).append(int(o["iters"][b]))
print("stragglers (tick: iters of its occurrences):", dict(sorted(hist.items())))
