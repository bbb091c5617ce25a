"""Per-instance iteration counts of the benchmark replay (tick t-1 and tick t) -> gpurun_out/iters_dump.npz (launch-order studies)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cmpc_loader
pkg = cmpc_loader.load()
N, B = 20, 4096
WM = int(sys.argv[1]) if len(sys.argv) > 1 else 4
(prev2, prev, cur), mass, k1, idx = bench.replay_workload(N, B, seed=0, back=2)
s = pkg.BatchSolver(N, B, device=0)
s.solve_host(*prev2, mass, k1, 0)
o1 = s.solve_host(*prev, mass, k1, WM)
o2 = s.solve_host(*cur, mass, k1, WM)
st = s.last_stats()
np.savez(os.path.join(ROOT, "gpurun_out", "iters_dump.npz"), idx=idx, it_prev=o1["iters"], it_cur=o2["iters"], st_cur=o2["status"], kernel_ms=st["kernel_ms"])
print("prev mean %.2f cur mean %.2f max %d kernel %.2f ms" % (o1["iters"].mean(), o2["iters"].mean(), o2["iters"].max(), st["kernel_ms"]))
