"""First GPU sanity/timing run: FP64 probe, batches of recorded walk ticks (cold / full warm / primal warm)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpc_loader

pkg = cmpc_loader.load()
print("fp64 peak TFLOP/s", pkg.measure_fp64_peak(0), flush=True)
for N in (10, 20):
    path = os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N)
    if not os.path.exists(path):
        continue
    w = np.load(path)
    T = len(w["x0"])
    rng = np.random.default_rng(0)
    for B in (256, 4096):
        idx = rng.integers(1, T, B)
        s = pkg.BatchSolver(N, B, device=0)
        arg = lambda ii: (w["x0"][ii], w["com_ref"][ii], w["foot_ref"][ii], w["gamma"][ii], float(w["mass"]), float(w["k1"]))
        for rep in range(2):
            t0 = time.time()
            out = s.solve_host(*arg(idx), 0)
            dt = time.time() - t0
        st = s.last_stats()
        print(json.dumps({"N": N, "B": B, "mode": "cold", "wall_ms": dt * 1e3, "kernel_ms": st["kernel_ms"],
                          "solves_per_s": B / (st["kernel_ms"] * 1e-3), "status": np.bincount(out["status"], minlength=6).tolist(),
                          "iters_mean": float(out["iters"].mean()), "iters_max": int(out["iters"].max()),
                          "nfact": st["nfact"], "nreg": st["nreg"], "viol_max": float(out["viol"].max())}), flush=True)
        # MPC tick: previous tick solved (state stays on the device), then the sampled tick warm-started from it
        for mode, name in ((2, "warm_full"), (1, "warm_primal")):
            s.solve_host(*arg(idx - 1), 0)
            out = s.solve_host(*arg(idx), mode)
            st = s.last_stats()
            print(json.dumps({"N": N, "B": B, "mode": name, "kernel_ms": st["kernel_ms"],
                              "solves_per_s": B / (st["kernel_ms"] * 1e-3), "status": np.bincount(out["status"], minlength=6).tolist(),
                              "iters_mean": float(out["iters"].mean()), "iters_max": int(out["iters"].max()),
                              "nfact": st["nfact"], "nreg": st["nreg"]}), flush=True)
        print("footprint", s.footprint(), flush=True)
        s.close()
