"""GPU parity against the round-2 oracle-T golden vectors (tests/golden/make_golden_r2.py): the payload variant
(k1 = 7 of `code/centroidal_mpc_vertices_payload.py:27-31`, per-instance masses 40.05 / 45 / 50 kg), perturbed initial
CoM / momentum states (BASELINE config 3 recipe) and the long horizon N = 60 (config 5).  Same bars as the nominal
golden ticks: cost 1e-6, x1 1e-6, u0 1e-4 (modulo the internal force), violation 1e-6 -- every instance, no
"up to 10 % may differ" allowance."""
import numpy as np
import pytest

from parity import COST_TOL, U0_TOL, VIOL_TOL, X1_TOL, golden_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,N", [("payload_N10", 10), ("perturbed_N20", 20), ("N60", 60)])
@pytest.mark.parametrize("warm", [0, 2])
def test_round2_golden_parity(pkg, golden, name, N, warm):
    g = golden[name]
    B = len(g["ticks"])
    s = pkg.BatchSolver(N, B, device=0)
    args = (g["x0"], g["com_ref"], g["foot_ref"], g["gamma"], g["mass"], g["k1"])
    out = s.solve_host(*args, 0)
    if warm:                                                   # re-solve from the converged iterate (full warm start)
        out = s.solve_host(*args, warm)
    assert (out["status"] == 0).all(), (name, out["status"])
    assert out["viol"].max() <= VIOL_TOL
    for k in range(B):
        ec, ex, eu = golden_errors(g, k, out["cost"][k], out["x1"][k], out["u0"][k])
        assert ec <= COST_TOL and ex <= X1_TOL and eu <= U0_TOL, (name, int(g["ticks"][k]), ec, ex, eu)
