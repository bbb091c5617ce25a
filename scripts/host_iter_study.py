"""CPU study of iteration counts (tests/hostsim build of the solver core): warm-started tick replay, parameter overrides."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import hostsim
from test_hostsim_parity import problem
from multiprocessing import Pool

N = int(os.environ.get("N", 20)); NS = int(os.environ.get("NS", 96))
w = dict(np.load(os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N)))
rng = np.random.default_rng(0)
idx = rng.integers(1, len(w["x0"]), NS)

def one(args):
    t, over, warm = args
    r0 = hostsim.solve(problem(w, t - 1, N))
    r = hostsim.solve(problem(w, t, N), work=r0["work"].copy(), warm=warm, **over)
    return (r0["status"], r["status"], r["iters"], r["nfact"], r["nreg"], r["cost"])

if __name__ == "__main__":
    hostsim.build()
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [{}]
    warm = int(os.environ.get("WARM", 2))
    base = None
    with Pool(8) as pool:
        for over in variants:
            t0 = time.time()
            res = np.array(pool.map(one, [(int(t), over, warm) for t in idx]))
            ok = res[:, 1] == 0
            if base is None: base = res[:, 5]
            dc = np.abs(res[:, 5] - base) / np.maximum(1, np.abs(base))
            print(over, "conv %d/%d iters %.2f nfact %.2f nreg %.2f p90it %d maxit %d max|dcost| %.1e  (%.0fs)" %
                  (ok.sum(), NS, res[ok, 2].mean(), res[ok, 3].mean(), res[ok, 4].mean(), np.percentile(res[ok, 2], 90), res[ok, 2].max(), dc[ok].max(), time.time() - t0), flush=True)
