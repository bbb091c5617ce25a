"""Cross-check EVERY oracle-T golden vector with scipy (trust-constr, then SLSQP where trust-constr does not return: standing
ticks, where LICQ fails) on the literal NLP with sympy derivatives, started 1e-3 away from oracle-T's answer
(oracle/scipy_check.py), and store the distances in tests/golden/scipy_pin.json.

    python tests/golden/make_scipy_pin.py            # all instances of golden_N10 / N20 / payload_N10 / perturbed_N20 (71) + three N = 60
                                                     # instances: about 40 CPU-minutes on 8 cores (the standing ticks take 4-6 minutes each)

tests/test_oracle_scipy.py asserts the stored distances against the parity tolerances and repeats two of the runs live."""
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FILES = [("N10", 10), ("N20", 20), ("payload_N10", 10), ("perturbed_N20", 20)]
N60_TICKS = (805, 1500, 1910)          # push window, mid walk, last valid tick: 1.5 - 2.5 minutes each (ticks 230 / 262 did not finish in 40)
# stored ALTERNATIVE KKT points (golden file, horizon, tick, index into X_alt / U_alt).  N = 60: the lower-cost points scipy moves to from the
# primary ones (6 / 10 minutes).  Not listed: N = 60 tick 1500 alternative 2 -- started next to it scipy goes to the primary point instead
ALTS = (("N20", 20, 255, 1), ("N20", 20, 759, 1), ("N60", 60, 150, 2), ("N60", 60, 230, 2))


def run(case):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle.scipy_check import cross_check
    from parity import COST_TOL, U0_TOL, X1_TOL, u0_err
    name, N, k = case[:3]
    alt = case[3] if len(case) > 3 else None
    g = np.load(os.path.join(HERE, "golden_%s.npz" % name))
    if "status" in g.files and g["status"][k] != 0:
        return None                                               # (the oracle itself did not converge there: not a golden point)
    per = np.ndim(g["mass"]) > 0
    mass, k1 = (float(g["mass"][k]), float(g["k1"][k])) if per else (float(g["mass"]), float(g["k1"]))
    out = None
    for method in ("trust-constr", "SLSQP"):
        Xs, Us = (g["X"][k], g["U"][k]) if alt is None else (g["X_alt"][k][alt], g["U_alt"][k][alt])
        r = cross_check(N, g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], mass, k1, Xs, Us, method=method, u0_metric=u0_err)
        r.update(golden=name, N=N, tick=int(g["ticks"][k]), index=int(k), mass=mass, k1=k1)
        if alt is not None:
            r.update(alt=int(alt))
        ok = r["cost_err"] <= COST_TOL and r["x1_err"] <= X1_TOL and r["u0_err"] <= U0_TOL and r["viol"] <= 1.1e-8
        if out is None or ok:
            out = r
        if ok:
            break
    print(json.dumps(out), flush=True)
    return out


def main():
    cases = []
    for name, N in FILES:
        g = np.load(os.path.join(HERE, "golden_%s.npz" % name))
        cases += [(name, N, k) for k in range(len(g["ticks"]))]
    g = np.load(os.path.join(HERE, "golden_N60.npz"))
    cases += [("N60", 60, list(g["ticks"]).index(t)) for t in N60_TICKS]
    for name, N, t, a in ALTS:
        cases.append((name, N, list(np.load(os.path.join(HERE, "golden_%s.npz" % name))["ticks"]).index(t), a))
    with Pool(8) as pool:
        res = [r for r in pool.map(run, cases, chunksize=1) if r is not None]
    with open(os.path.join(HERE, "scipy_pin.json"), "w") as f:
        json.dump(res, f, indent=0)


if __name__ == "__main__":
    main()
