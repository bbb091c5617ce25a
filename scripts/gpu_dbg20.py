import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cmpc_loader
from parity import cost_err, u0_err
pkg = cmpc_loader.load()
g = dict(np.load(os.path.join(ROOT, "tests/golden/golden_N20.npz")))
B = len(g["ticks"])
for rep in range(3):
    s = pkg.BatchSolver(20, B, device=0)
    out = s.solve_host(g["x0"], g["com_ref"], g["foot_ref"], g["gamma"], float(g["mass"]), float(g["k1"]), 0)
    ce = cost_err(out["cost"], g["cost"]); xe = np.abs(out["x1"][:, :12] - g["X"][:, 1, :12]).max(axis=1)
    ue = u0_err(out["u0"], g["U"][:, 0], g["x0"], g["gamma"][:, 0])
    for k in range(B):
        print(rep, int(g["ticks"][k]), "status", out["status"][k], "it", out["iters"][k], "cost_err %.2e x1 %.2e u0 %.2e viol %.1e" % (ce[k], xe[k], ue[k], out["viol"][k]))
