"""Cross-check oracle-T golden vectors with scipy (trust-constr / SLSQP on the literal NLP with sympy derivatives, started
1e-3 away from oracle-T's answer; oracle/scipy_check.py) and store the distances in tests/golden/scipy_pin.json.

    python tests/golden/make_scipy_pin.py

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

# (golden file, horizon, tick, method): standing ticks -> SLSQP (trust-constr stalls where LICQ fails), others trust-constr
CASES = [("N10", 10, 0, "SLSQP"), ("N10", 10, 200, "trust-constr"), ("N10", 10, 262, "trust-constr"), ("N10", 10, 805, "trust-constr"),
         ("N10", 10, 1960, "trust-constr"), ("N20", 20, 255, "trust-constr"), ("N20", 20, 850, "trust-constr"),
         ("payload_N10", 10, 262, "trust-constr"), ("payload_N10", 10, 805, "trust-constr"), ("payload_N10", 10, 150, "SLSQP"),
         ("perturbed_N20", 20, None, "trust-constr"), ("perturbed_N20", 20, None, "trust-constr")]


def run(case):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle.scipy_check import cross_check
    from parity import u0_err
    name, N, tick, method, k = case
    g = np.load(os.path.join(HERE, "golden_%s.npz" % name))
    per = np.ndim(g["mass"]) > 0
    mass, k1 = (float(g["mass"][k]), float(g["k1"][k])) if per else (float(g["mass"]), float(g["k1"]))
    r = cross_check(N, g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], mass, k1, g["X"][k], g["U"][k], method=method,
                    u0_metric=u0_err)
    r.update(golden=name, N=N, tick=int(g["ticks"][k]), index=int(k), mass=mass, k1=k1)
    print(json.dumps(r), flush=True)
    return r


def main():
    cases = []
    nper = 0
    for name, N, tick, method in CASES:
        g = np.load(os.path.join(HERE, "golden_%s.npz" % name))
        if tick is None:
            k = nper; nper += 1
        else:
            k = list(g["ticks"]).index(tick)
        cases.append((name, N, tick, method, k))
    with Pool(6) as pool:
        res = pool.map(run, cases, chunksize=1)
    with open(os.path.join(HERE, "scipy_pin.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
