import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import cmpc_loader; pkg = cmpc_loader.load()
from oracle.walk import load_walk
from cmpc_b200.assembly import PlanTables, ReferenceTables, assemble_tick, pack_instances
N, B = 60, 2048
planner, com_ref, params, initial = load_walk(); params = dict(params, N=N)
tables, refs = PlanTables(planner.plan), ReferenceTables(com_ref, planner)
w = dict(np.load("tests/golden/walk_ticks_N20.npz"))
rng = np.random.default_rng(3); ticks = rng.integers(0, 1900, B)
def instance(t):
    x = w["x0"][t]
    cur = {"com": {"pos": x[0:3], "vel": x[3:6]}, "hw": {"val": x[6:9]}, "lfoot": {"pos": [0, 0, x[12]]}, "rfoot": {"pos": [0, 0, x[16]]}}
    return assemble_tick(tables, refs, planner.plan, params, cur, x[9:12], int(t))
now = pack_instances([instance(t) for t in ticks])
for over in ({}, {"max_iter": 300}, {"eps_reg": 1e-9}, {"eps_reg": 1e-4}):
    s = pkg.BatchSolver(N, B, device=0, **over)
    out = s.solve_host(*now, float(w["mass"]), float(w["k1"]), 0); st = s.last_stats()
    print(over, "conv", (out["status"] == 0).mean(), "hist", np.bincount(out["status"], minlength=6).tolist(), "iters/solve", st["iters"] / B, "ms", st["kernel_ms"], flush=True)
    s.close()
