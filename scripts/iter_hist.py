import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpc_loader
pkg = cmpc_loader.load()
N, B = 20, 4096
w = np.load(os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N))
rng = np.random.default_rng(0)
idx = rng.integers(1, len(w["x0"]), B)
arg = lambda ii: (w["x0"][ii], w["com_ref"][ii], w["foot_ref"][ii], w["gamma"][ii], float(w["mass"]), float(w["k1"]))
for over in ({}, {"max_iter": 60}, {"mu_warm": 1e-5}, {"ls_max": 1}, {"ls_max": 0}):
    s = pkg.BatchSolver(N, B, device=0, **over)
    s.solve_host(*arg(idx - 1), 0)
    out = s.solve_host(*arg(idx), 2)
    st = s.last_stats()
    it = out["iters"]
    print(over, "kernel_ms %.1f solves/s %.0f conv %d iters mean %.2f p50 %d p90 %d p99 %d max %d nfact %d" %
          (st["kernel_ms"], B / st["kernel_ms"] * 1e3, (out["status"] == 0).sum(), it.mean(), np.percentile(it, 50), np.percentile(it, 90),
           np.percentile(it, 99), it.max(), st["nfact"]), flush=True)
    s.close()
