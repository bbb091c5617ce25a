"""Per-phase cycle breakdown of the solve kernel (library built with -DCMPC_PROFILE)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmpc_loader
pkg = cmpc_loader.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
w = np.load(os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N))
rng = np.random.default_rng(0)
idx = rng.integers(1, len(w["x0"]), B)
s = pkg.BatchSolver(N, B, device=0)
arg = lambda ii: (w["x0"][ii], w["com_ref"][ii], w["foot_ref"][ii], w["gamma"][ii], float(w["mass"]), float(w["k1"]))
for mode in (0, 2):
    if mode == 2:
        s.solve_host(*arg(idx - 1), 0)
    out = s.solve_host(*arg(idx), mode)
    st = s.last_stats(); pc = s.phase_cycles()
    solve_total, cta_total = pc.pop("solve_total", 0), pc.pop("cta_total", 0); tot = sum(pc.values()) or 1
    print(json.dumps({"N": N, "B": B, "mode": mode, "kernel_ms": st["kernel_ms"], "solves_per_s": B / st["kernel_ms"] * 1e3,
                      "iters": st["iters"], "nfact": st["nfact"], "phases_over_solve": round(tot / max(solve_total, 1), 3),
                      "solve_over_cta": round(solve_total / max(cta_total, 1), 3), "cta_cycles_mean": cta_total / min(B, s.footprint()["slots"]), "kernel_cycles": st["kernel_ms"] * 1.965e6, "conv": int((out["status"] == 0).sum()),
                      "cycles_per_fact": {k: round(v / max(st["nfact"], 1)) for k, v in pc.items()},
                      "share": {k: round(v / tot, 3) for k, v in pc.items()}}), flush=True)
