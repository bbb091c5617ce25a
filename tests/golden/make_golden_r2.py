"""Generate the round-2 golden vectors (ORACLE-T = oracle/ipm_py.py, literal NLP, sympy derivatives, dense LDL'):

  golden_payload_N10.npz    payload variant: k1 = 7 (`code/centroidal_mpc_vertices_payload.py:27-31`), per-instance
                            mass in {40.05487735, 45, 50}, 12 ticks of the recorded walk (standing, single support,
                            landing, push window)
  golden_perturbed_N20.npz  BASELINE config 3 recipe: base ticks of the recorded N = 20 walk (seed 1),
                            x0 perturbed (CoM 1 cm, velocity 5 cm/s, angular momentum from the cuhw.txt scale, theta_hat
                            2 N); the first 20 instances oracle-T converges on
  golden_N60.npz            BASELINE config 5: horizon N = 60, ticks 0 (standing), 150 (first lift-off inside), 230
                            (single support), 262 (landings inside), 805 (push window), 1500, 1910 (last valid tick)

Every instance is solved from the neutral start (x_i = x0, f_z = m g / #contact vertices) for three initial barrier
values; all converged solutions are stored (cost_alt / X_alt / U_alt): the NLP is non-convex and parity means "equal to
one of the oracle's KKT points" (DESIGN.md section 3).

    python tests/golden/make_golden_r2.py [payload] [perturbed] [n60]
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
MU0 = [0.1, 0.11, 1.0]
MASS0 = 40.05487735


def run(job):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ipm_py
    N, args, mu0, max_iter = job
    prob = ipm_py.unpack_problem(N, *args)
    o = ipm_py.Options.oracle_T()
    o.mu_init, o.max_iter = mu0, max_iter
    t0 = time.time()
    r = ipm_py.solve(prob, ipm_py.neutral_start(prob), o)
    return r.status, r.cost, r.viol, r.iters, r.X(N).T, r.U(N).T, time.time() - t0


def solve_all(N, x0, com, foot, gam, mass, k1, max_iter=400, procs=8):
    B = len(x0)
    jobs = [(N, (x0[k], com[k], foot[k], gam[k], float(mass[k]), float(k1[k])), mu0, max_iter) for mu0 in MU0 for k in range(B)]
    with Pool(procs) as pool:
        res = pool.map(run, jobs, chunksize=1)
    A = len(MU0)
    cost = np.full((B, A), np.nan); viol = np.full((B, A), np.nan); iters = np.zeros((B, A), int)
    X = np.full((B, A, N + 1, 20), np.nan); U = np.full((B, A, N, 32), np.nan)
    for a in range(A):
        for k in range(B):
            st, c, v, it, x, u, dt = res[a * B + k]
            iters[k, a] = it
            if st == 0:
                cost[k, a], viol[k, a], X[k, a], U[k, a] = c, v, x, u
            print("  inst %2d mu0 %.2f status %d iters %3d cost %.9e viol %.1e (%.0f s)" % (k, MU0[a], st, it, c, v, dt), flush=True)
    return cost, viol, iters, X, U


def save(name, N, ticks, x0, com, foot, gam, mass, k1, out, keep=None):
    cost, viol, iters, X, U = out
    ok = np.isfinite(cost).any(axis=1)
    sel = np.flatnonzero(ok) if keep is None else np.flatnonzero(ok)[:keep]
    first = np.array([np.flatnonzero(np.isfinite(cost[k]))[0] for k in sel])
    np.savez_compressed(
        os.path.join(HERE, name), ticks=np.asarray(ticks)[sel], x0=x0[sel], com_ref=com[sel], foot_ref=foot[sel], gamma=gam[sel],
        mass=np.asarray(mass, float)[sel], k1=np.asarray(k1, float)[sel], status=np.zeros(len(sel), int),
        iters=iters[sel, first], cost=cost[sel, first], viol=viol[sel, first], X=X[sel, first], U=U[sel, first],
        cost_alt=cost[sel], X_alt=X[sel], U_alt=U[sel])
    d = np.nanmax(np.abs(cost[sel] - cost[sel, first][:, None]) / np.maximum(1, np.abs(cost[sel, first][:, None])), axis=1)
    print("%s: %d instances stored (%d solved); with a second KKT point: %s" %
          (name, len(sel), int(ok.sum()), [(int(np.asarray(ticks)[k]), float(dd)) for k, dd in zip(sel, d) if dd > 1e-7]), flush=True)


def payload():
    N = 10
    w = dict(np.load(os.path.join(HERE, "walk_ticks_N10.npz")))
    ticks = np.array([0, 150, 199, 200, 230, 262, 270, 300, 805, 850, 1275, 1775])
    mass = np.array([MASS0, 45.0, 50.0] * 4)
    k1 = np.full(len(ticks), 7.0)
    args = (w["x0"][ticks], w["com_ref"][ticks], w["foot_ref"][ticks], w["gamma"][ticks])
    save("golden_payload_N10.npz", N, ticks, *args, mass, k1, solve_all(N, *args, mass, k1))


def perturbed():
    N, B = 20, 28
    w = dict(np.load(os.path.join(HERE, "walk_ticks_N20.npz")))
    rng = np.random.default_rng(1)
    idx = rng.integers(0, len(w["x0"]), B)
    x0 = w["x0"][idx].copy()                                  # SURVEY.md 8d recipe (same as tests/test_gpu_configs_full.py)
    x0[:, 0:3] += rng.normal(0, 0.01, (B, 3)); x0[:, 2] = np.minimum(x0[:, 2], 0.759)
    x0[:, 3:6] += rng.normal(0, 0.05, (B, 3))
    x0[:, 6:9] = rng.normal(0, 1.0, (B, 3)) * np.array([0.88, 0.63, 0.20])
    x0[:, 9:12] = rng.normal(0, 2.0, (B, 3))
    mass, k1 = np.full(B, float(w["mass"])), np.full(B, float(w["k1"]))
    args = (x0, w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx])
    save("golden_perturbed_N20.npz", N, idx, *args, mass, k1, solve_all(N, *args, mass, k1, max_iter=300), keep=20)


def n60():
    from oracle.walk import load_walk
    import cmpc_loader
    cmpc_loader.load()
    from cmpc_b200 import assembly as asm
    N = 60
    planner, com_ref, params, initial = load_walk()
    params = dict(params, N=N)
    tables, refs = asm.PlanTables(planner.plan), asm.ReferenceTables(com_ref, planner)
    w = dict(np.load(os.path.join(HERE, "walk_ticks_N20.npz")))
    ticks = np.array([0, 150, 230, 262, 805, 1500, 1910])

    def instance(t):
        x = w["x0"][t]                                         # state the recorded (N = 20) walk had at tick t
        cur = {"com": {"pos": x[0:3], "vel": x[3:6]}, "hw": {"val": x[6:9]}, "lfoot": {"pos": [0, 0, x[12]]}, "rfoot": {"pos": [0, 0, x[16]]}}
        return asm.assemble_tick(tables, refs, planner.plan, params, cur, x[9:12], int(t))

    args = asm.pack_instances([instance(t) for t in ticks])
    mass, k1 = np.full(len(ticks), float(w["mass"])), np.full(len(ticks), float(w["k1"]))
    save("golden_N60.npz", N, ticks, *args, mass, k1, solve_all(N, *args, mass, k1, max_iter=500, procs=int(os.environ.get("PROCS", 7))))


if __name__ == "__main__":
    what = sys.argv[1:] or ["payload", "perturbed", "n60"]
    for name in what:
        t0 = time.time()
        {"payload": payload, "perturbed": perturbed, "n60": n60}[name]()
        print("%s done in %.0f s" % (name, time.time() - t0), flush=True)
