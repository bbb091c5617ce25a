"""Independent pin of the oracle: scipy's trust-constr / SLSQP on the LITERAL NLP
(oracle/spec.py, sympy derivatives), started 1e-3 away from oracle-T's answer, return to that answer within the parity
tolerances.  tests/golden/scipy_pin.json holds the distances of EVERY golden instance at N = 10 / 20 (25 + 14 nominal ticks,
12 payload instances with k1 = 7 and three masses, 20 perturbed initial states) and of three N = 60 instances (push window,
mid walk, last valid tick), produced by tests/golden/make_scipy_pin.py; two instances are repeated live here."""
import json
import os

import numpy as np
import pytest

from parity import COST_TOL, U0_TOL, X1_TOL, u0_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_scipy_cross_checks_are_within_the_parity_tolerances():
    pin = json.load(open(os.path.join(ROOT, "tests", "golden", "scipy_pin.json")))
    assert len(pin) >= 74
    kinds = {(p["golden"], p["method"]) for p in pin}
    assert ("payload_N10", "trust-constr") in kinds and ("perturbed_N20", "trust-constr") in kinds and ("N10", "SLSQP") in kinds
    assert any(p["k1"] == 7.0 and p["mass"] > 44 for p in pin)
    assert sum(1 for p in pin if p["N"] == 60) >= 3                     # the long horizon is pinned too
    # every golden instance of the four N = 10 / 20 files is there
    for name in ("N10", "N20", "payload_N10", "perturbed_N20"):
        g = np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % name))
        conv = [k for k in range(len(g["ticks"])) if "status" not in g.files or g["status"][k] == 0]
        assert sorted(p["index"] for p in pin if p["golden"] == name and p.get("alt") is None) == conv, name
    assert sum(1 for p in pin if p.get("alt") is not None) >= 4        # stored alternative KKT points (N = 20: two, N = 60: two)
    for p in pin:
        assert p["start_dist"] >= 1e-3                                  # it did start away from the answer
        assert p["cost_err"] <= COST_TOL and p["x1_err"] <= X1_TOL and p["u0_err"] <= U0_TOL, p
        assert p["viol"] <= 1.1e-8                                      # feasible for the relaxed rows (bound_relax_factor 1e-8)


@pytest.mark.parametrize("name,N,tick,method", [("N10", 10, 1960, "trust-constr"), ("N10", 10, 200, "trust-constr")])
def test_scipy_returns_to_the_oracle_point_live(golden, name, N, tick, method):
    from oracle.scipy_check import cross_check
    g = golden[N]
    k = list(g["ticks"]).index(tick)
    r = cross_check(N, g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], float(g["mass"]), float(g["k1"]), g["X"][k], g["U"][k],
                    method=method, u0_metric=u0_err)
    assert r["cost_err"] <= COST_TOL and r["x1_err"] <= X1_TOL and r["u0_err"] <= U0_TOL, r
