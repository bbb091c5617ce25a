"""CPU tests of the oracle itself and of the host-side logic of the product.

  * the C oracle (oracle/ipm_c.c, sympy-generated derivatives) reproduces the golden vectors produced by the
    independent numpy/scipy oracle (oracle/ipm_py.py, generic dense-LDL' KKT solves);
  * the numpy oracle reproduces its own golden vector (guards the fixtures against code drift);
  * k2 cancels out of the reference's Lyapunov row, symbolically;
  * the product's vectorised parameter assembly equals the literal restatement of MPC file :482-600 on every tick
    class (t < 200, single support, landing, last valid tick) including the column-major yaw quirk.
"""
import numpy as np
import pytest
import sympy as sp

from parity import COST_TOL, U0_TOL, X1_TOL, golden_errors


@pytest.mark.parametrize("N", [10, 20])
def test_c_oracle_reproduces_golden(golden, N):
    from oracle import ipm_c
    g = golden[N]
    bad = 0
    for k in range(len(g["ticks"])):
        r = ipm_c.solve_packed(N, g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], float(g["mass"]), float(g["k1"]), max_iter=150)
        if r["status"] != 0:
            bad += 1
            continue
        ec, ex, eu = golden_errors(g, k, r["cost"], r["x1"], r["u0"])
        assert ec <= COST_TOL and ex <= X1_TOL and eu <= U0_TOL, (N, int(g["ticks"][k]), ec, ex, eu)
        assert r["viol"] <= 1e-6
    assert bad <= 1          # N=20 tick 759 sits between two KKT points 3e-6 apart and can run out of iterations


def test_numpy_oracle_reproduces_its_golden_vector(golden):
    from oracle import ipm_py
    g = golden[10]
    k = list(g["ticks"]).index(805)
    r = ipm_py.solve_packed(10, g["x0"][k], g["com_ref"][k], g["foot_ref"][k], g["gamma"][k], float(g["mass"]), float(g["k1"]))
    assert r["status"] == 0
    ec, ex, eu = golden_errors(g, k, r["cost"], r["x1"], r["u0"])
    assert ec <= 1e-9 and ex <= 1e-8 and eu <= 1e-7


def test_k2_cancels_symbolically():
    from oracle.spec import build_block
    V, Pv, d, g, cost = build_block()
    k2 = [s for s in Pv if s.name == "k2"][0]
    assert sp.simplify(sp.diff(g[0], k2)) == 0           # the Lyapunov row does not depend on k2
    assert sp.diff(cost, k2) == 0 and all(sp.diff(e, k2) == 0 for e in d)


def test_generated_c_matches_sympy_block_functions():
    """stage_gen.c (stage-wise form) against the lambdified literal block: dynamics and Lyapunov row at a random point."""
    import ctypes
    from oracle import ipm_c
    from oracle.spec import block_functions, PARAM_NAMES
    L = ipm_c.lib()
    rng = np.random.default_rng(3)
    z = rng.normal(size=60); z[2 + 32] = 0.7
    p = np.zeros(40); p[0:9] = rng.normal(size=9) * 0.1; p[26] = 1.0; p[27] = 0.0; p[31] = 1; p[33] = 1
    p[34], p[35], p[36], p[37], p[38], p[39] = 1.0, 40.0, 4.0, 0.01, 9.81, 1e-9
    out = {n: np.zeros(k) for n, k in (("cost", 1), ("grad", 60), ("phi", 28), ("jphi", 200), ("g", 55), ("jg", 200), ("hess", 900))}
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    L.sg_stage(dp(z), dp(p), dp(np.zeros(28)), dp(np.zeros(55)), dp(out["cost"]), dp(out["grad"]), dp(out["phi"]), dp(out["jphi"]),
               dp(out["g"]), dp(out["jg"]), dp(out["hess"]))
    bf = block_functions()
    v = np.zeros(104); v[0:20] = z[32:52]; v[20:52] = z[0:32]; v[52:72] = out["phi"][:20]
    P = np.zeros(len(PARAM_NAMES)); pi = {n: j for j, n in enumerate(PARAM_NAMES)}
    P[0:9] = p[0:9]; P[pi["gl"]], P[pi["gr"]] = 1.0, 0.0; P[pi["hw_on"]] = 1.0
    P[pi["mass"]], P[pi["k1"]], P[pi["k2"]], P[pi["delta"]], P[pi["grav"]] = 40.0, 4.0, 0.37, 0.01, 9.81
    dres = np.array([float(e) for e in bf.f_d(list(v), list(P))])
    assert np.abs(dres).max() < 1e-12                    # x_{i+1} = phi(x_i, u_i) satisfies the literal defect
    glit = np.array([float(e) for e in bf.f_g(list(v), list(P))])
    assert abs(glit[0] - out["g"][0]) < 1e-9 and abs(glit[1] - out["g"][1]) < 1e-9       # Lyapunov / angular momentum rows
    assert np.abs(glit[3:19] - out["g"][3:19]).max() < 1e-12        # friction rows of the foot in contact (gamma_l = 1)


@pytest.mark.parametrize("t", [0, 57, 199, 200, 262, 270, 1500, 1950])
def test_vectorised_assembly_matches_literal_restatement(pkg, t):
    from oracle.walk import assemble, load_walk
    from cmpc_b200.assembly import PlanTables, ReferenceTables, assemble_tick
    planner, com_ref, params, initial = load_walk()
    params["N"] = 20
    # make the yaw reference non-trivial so the column-major quirk (:599-600) is exercised
    planner.position_contacts_ref["contact_left"][:, 2] = 0.001 * np.arange(len(planner.position_contacts_ref["contact_left"]))
    planner.position_contacts_ref["contact_right"][:, 2] = -0.002 * np.arange(len(planner.position_contacts_ref["contact_right"]))
    cur = {"com": {"pos": np.array([0.1, 0.02, 0.71]), "vel": np.array([0.1, -0.1, 0.0])}, "hw": {"val": np.array([0.3, -0.2, 0.05])},
           "lfoot": {"pos": np.array([0, 0, 0.01, 0.2, 0.1, 0.0])}, "rfoot": {"pos": np.array([0, 0, -0.02, 0.1, -0.1, 0.0])}}
    th = np.array([0.5, -0.4, 0.3])
    prob = assemble(planner, com_ref, params, cur, th, t)
    x0, com, foot, gam = assemble_tick(PlanTables(planner.plan), ReferenceTables(com_ref, planner), planner.plan, params, cur, th, t)
    assert np.array_equal(x0, prob.x0) and np.array_equal(com.T, prob.com_ref)
    assert np.array_equal(foot[:, 0:3].T, prob.pl_ref) and np.array_equal(foot[:, 3:6].T, prob.pr_ref)
    assert np.array_equal(foot[:, 6], prob.al_ref) and np.array_equal(foot[:, 7], prob.ar_ref)
    assert np.array_equal(gam[:, 0], prob.gl) and np.array_equal(gam[:, 1], prob.gr)


def test_plan_tables_match_planner_queries(pkg):
    from oracle.walk import load_walk
    from cmpc_b200.assembly import PlanTables
    planner, _, _, _ = load_walk()
    tb = PlanTables(planner.plan)
    for t in list(range(0, 2000, 7)) + [199, 200, 269, 270, 299, 300, 1999, 2099]:
        assert tb.step_index_at(t) == planner.get_step_index_at_time(t)
        assert tb.phase_at(t) == planner.get_phase_at_time(t)
    with pytest.raises((TypeError, IndexError)):
        tb.phase_at(2100)                                 # beyond the 20-step plan: the reference dies here too (step index None)
    with pytest.raises(IndexError):
        from cmpc_b200.assembly import ReferenceTables, assemble_tick
        planner2, com_ref, params, initial = load_walk()
        params["N"] = 10
        cur = {"com": {"pos": np.zeros(3), "vel": np.zeros(3)}, "hw": {"val": np.zeros(3)}, "lfoot": {"pos": np.zeros(6)}, "rfoot": {"pos": np.zeros(6)}}
        assemble_tick(PlanTables(planner2.plan), ReferenceTables(com_ref, planner2), planner2.plan, params, cur, np.zeros(3), 1961)
