"""Per instance sample of the bench replay (seeds 0..3): kernel time, the longest solves (iterations, tick) -- which instance is the straggler?
usage: python scripts/straggler_probe.py lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import bench, cmpc_loader
pkg = cmpc_loader.load()
N, B = 20, 4096
for seed in range(4):
    (prev2, prev, cur), mass, k1, idx = bench.replay_workload(N, B, seed=seed, back=2, ahead=0, headroom=32)
    s = pkg.BatchSolver(N, B, device=0)
    s.solve_host(*prev2, mass, k1, 0)
    s.solve_host(*prev, mass, k1, 4)
    s.warm_save(B)
    s.warm_restore(B)
    o = s.solve_host(*cur, mass, k1, 4)
    st = s.last_stats()
    it = o["iters"]; top = np.argsort(-it)[:6]
    print(json.dumps({"seed": seed, "kernel_ms": round(st["kernel_ms"], 2), "iters": round(float(it.mean()), 3), "top": [(int(idx[k]), int(it[k])) for k in top], "n_ge40": int((it >= 40).sum())}), flush=True)
''' % ROOT
for so in sys.argv[1:]:
    print("==", so, flush=True)
    env = dict(os.environ, CMPC_LIB=os.path.join(ROOT, so))
    try:
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=200)
        print(out.stdout.strip() or out.stderr[-600:], flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT", flush=True)
