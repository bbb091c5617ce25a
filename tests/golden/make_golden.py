"""Generate tests/golden/golden_N{10,20}.npz : recorded-walk instances with ORACLE-T solutions.

Run in the development container:   python tests/golden/make_golden.py [--ticks-only]

1. Inputs.  The surrogate closed-loop walk of SURVEY.md section 8d (plant = centroidal model, measured
   angular momentum replayed from the reference's own recording `original_code/cuhw.txt`, lateral push for
   800 < t < 900) is driven once per horizon; every tick's (x0, com_ref, foot_ref, gamma) is recorded.  The
   loop is driven with the g++ host build of the product core (tests/hostsim, a test aid) because it is fast;
   the recorded INPUTS do not depend on which solver produced them beyond closing the loop.
2. Expected outputs.  A selection of ticks (standing, first single support, landings inside the horizon, the
   step-adjustment tick, push window, last ticks) is solved by the independent oracle (oracle/ipm_py.py,
   sympy derivatives + dense LDL' KKT solves, tolerance 1e-8 at mu = 1e-9), started from a neutral guess
   (x_i = x0, vertex f_z = m g / #contact vertices).  Stored: cost, u0, x1, xN, violation, iterations, and the
   full primal trajectory.
Also stores `walk_ticks_N*.npz`: inputs of EVERY tick (float64) for the replay batches of bench.py/tests.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))

TICKS = {10: [0, 1, 57, 108, 150, 191, 199, 200, 216, 230, 262, 269, 270, 285, 300, 473, 671, 805, 850, 891, 1275, 1500, 1775, 1900, 1960],
         20: [0, 150, 185, 200, 230, 255, 270, 300, 358, 562, 759, 850, 1455, 1950]}


def neutral_start(prob):
    N = prob.N
    w = np.zeros(52 * N + 20)
    for i in range(N + 1):
        w[52 * i:52 * i + 20] = prob.x0
    for i in range(N):
        n = prob.gl[i] + prob.gr[i]
        for v in range(8):
            ge = prob.gl[i] if v < 4 else prob.gr[i]
            w[52 * i + 20 + 3 * v + 2] = ge * prob.mass * prob.grav / (4 * max(n, 1))
    return w


def run_walk(N, t_end):
    import hostsim
    from oracle.mpc_ref import centroidal_mpc, surrogate_walk
    from oracle.walk import load_walk
    planner, com_ref, params, initial = load_walk()
    params["N"] = N
    state = {"work": None, "first": True}

    class R:
        pass

    def solver(prob, w0, opts):
        r = hostsim.solve(prob, work=state["work"], warm=0 if state["first"] else 2)
        state["first"] = False
        state["work"] = r["work"]
        res = R()
        res.status, res.iters, res.cost, res.viol = r["status"], r["iters"], r["cost"], r["viol"]
        X, U = r["X"][:20], r["U"]
        res.X = lambda n: X
        res.U = lambda n: U
        res.w = None
        return res

    mpc = centroidal_mpc(initial, planner, params, com_ref, solver=solver)
    mpc.use_warm = False
    rec = []
    surrogate_walk(mpc, initial, 0, t_end, params["mass"], record=lambda t, m, cur: rec.append((t, m.last_problem)),
                   hw_trace=initial["hw_meas"])
    return rec


def solve_oracle(prob):
    from oracle import ipm_py
    o = ipm_py.Options.oracle_T()
    o.max_iter = 400
    t0 = time.time()
    r = ipm_py.solve(prob, neutral_start(prob), o)
    return r, time.time() - t0


def pack(probs):
    import hostsim
    xs, cs, fs, gs = [], [], [], []
    for p in probs:
        x0, com, foot, gam = hostsim.pack(p)
        xs.append(x0); cs.append(com); fs.append(foot); gs.append(gam)
    return np.stack(xs), np.stack(cs), np.stack(fs), np.stack(gs)


def main():
    for N, t_end in ((10, 1961), (20, 1951)):
        t0 = time.time()
        rec = run_walk(N, t_end)
        print("N=%d walk: %d ticks in %.0f s" % (N, len(rec), time.time() - t0), flush=True)
        x0, com, foot, gam = pack([p for _, p in rec])
        np.savez_compressed(os.path.join(HERE, "walk_ticks_N%d.npz" % N), x0=x0, com_ref=com, foot_ref=foot, gamma=gam,
                            mass=rec[0][1].mass, k1=rec[0][1].k1)
        if "--ticks-only" in sys.argv:
            continue
        sel = [p for t, p in rec if t in TICKS[N]]
        with Pool(8) as pool:
            out = pool.map(solve_oracle, sel)
        for t, (r, dt) in zip(TICKS[N], out):
            print("  t=%4d status %d iters %3d cost %.9e viol %.2e  (%.1f s)" % (t, r.status, r.iters, r.cost, r.viol, dt), flush=True)
        sx0, scom, sfoot, sgam = pack(sel)
        np.savez_compressed(
            os.path.join(HERE, "golden_N%d.npz" % N), ticks=np.array(TICKS[N]), x0=sx0, com_ref=scom, foot_ref=sfoot, gamma=sgam,
            mass=sel[0].mass, k1=sel[0].k1,
            status=np.array([r.status for r, _ in out]), iters=np.array([r.iters for r, _ in out]),
            cost=np.array([r.cost for r, _ in out]), viol=np.array([r.viol for r, _ in out]),
            X=np.stack([r.X(N).T for r, _ in out]), U=np.stack([r.U(N).T for r, _ in out]))


if __name__ == "__main__":
    main()
