"""SM time per factorisation-iteration of several builds (-DCMPC_PROFILE): independent of the makespan tail, which moves with any
change of the rounding (another instance becomes the straggler).  usage: python scripts/prof_ab.py lib1.so lib2.so ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import bench, cmpc_loader
pkg = cmpc_loader.load()
N, B = 20, 4096
(prev2, prev, cur), mass, k1, idx = bench.replay_workload(N, B, seed=0, back=2)
s = pkg.BatchSolver(N, B, device=0)
s.solve_host(*prev2, mass, k1, 0)
s.solve_host(*prev, mass, k1, 4)
s.warm_save(B)
res = []
for rep in range(3):
    s.warm_restore(B)
    o = s.solve_host(*cur, mass, k1, 4)
    st = s.last_stats(); pc = s.phase_cycles()
    slots = min(B, s.footprint()["slots"])
    res.append((pc["solve_total"] / st["nfact"], st["kernel_ms"], st["nfact"] / B, int(o["iters"].max()), pc["solve_total"] / (slots * st["kernel_ms"] * 1.965e6)))
pc.pop("cta_total"); tot = pc.pop("solve_total")
print(json.dumps({"cycles_per_fact": float(np.median([r[0] for r in res])), "kernel_ms": float(np.median([r[1] for r in res])), "nfact": res[0][2], "max_iters": res[0][3], "slot_busy": round(float(np.median([r[4] for r in res])), 3),
                  "phases": {k: round(v / st["nfact"]) for k, v in pc.items()}}))
''' % ROOT
for so in sys.argv[1:]:
    env = dict(os.environ, CMPC_LIB=os.path.join(ROOT, so))
    try:                                            # a kernel that hangs must not eat the GPU budget
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=90)
        print(so, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
    except subprocess.TimeoutExpired:
        print(so, "TIMEOUT (hang?)", flush=True)
