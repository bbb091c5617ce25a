"""Add alternative oracle solutions to tests/golden/golden_N*.npz.

The NLP is non-convex (bilinear torque, indefinite Lyapunov rows): some instances have several KKT points a few
1e-6 apart in relative cost (e.g. N=20, tick 759: J = 83582.37033 and 83582.64861, both reached by the oracle
depending on the initial barrier parameter).  Parity is therefore "equal to ONE of the oracle's solutions of the
instance".  This script re-solves every golden instance with the oracle from the same neutral start for other
initial barrier values and stores all solutions as cost_alt / X_alt / U_alt (variant 0 = the original).
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
MU0 = [0.11, 1.0]


def run(job):
    from oracle import ipm_py
    N, args, mu0 = job
    prob = ipm_py.unpack_problem(N, *args)
    o = ipm_py.Options.oracle_T()
    o.mu_init, o.max_iter = mu0, 400
    r = ipm_py.solve(prob, ipm_py.neutral_start(prob), o)
    return r.status, r.cost, r.X(N).T, r.U(N).T


def main():
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    for N in (10, 20):
        path = os.path.join(HERE, "golden_N%d.npz" % N)
        g = dict(np.load(path))
        B = len(g["ticks"])
        jobs = [(N, (g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], float(g["mass"]), float(g["k1"])), mu0)
                for mu0 in MU0 for k in range(B)]
        with Pool(8) as pool:
            res = pool.map(run, jobs, chunksize=1)
        A = 1 + len(MU0)
        cost = np.full((B, A), np.nan); X = np.full((B, A) + g["X"].shape[1:], np.nan); U = np.full((B, A) + g["U"].shape[1:], np.nan)
        cost[:, 0], X[:, 0], U[:, 0] = g["cost"], g["X"], g["U"]
        for a in range(len(MU0)):
            for k in range(B):
                st, c, x, u = res[a * B + k]
                if st == 0:
                    cost[k, a + 1], X[k, a + 1], U[k, a + 1] = c, x, u
        g["cost_alt"], g["X_alt"], g["U_alt"] = cost, X, U
        np.savez_compressed(path, **g)
        d = np.nanmax(np.abs(cost - cost[:, :1]) / np.maximum(1, np.abs(cost[:, :1])), axis=1)
        print("N=%d: instances with a second KKT point:" % N, [(int(g["ticks"][k]), float(d[k])) for k in range(B) if d[k] > 1e-7], flush=True)


if __name__ == "__main__":
    main()
