"""CPU check of the product's solver CORE (csrc/cmpc_solver.h, cmpc_model.h) compiled by g++ with a serial
execution policy (tests/hostsim -- a test aid, never a product path) against the oracle's golden vectors.
Catches errors in the analytic derivatives / Riccati algebra here, where there is no GPU; the CUDA build of the
same source is checked by tests/test_gpu_parity.py on the B200."""
import ctypes
import os

import numpy as np
import pytest

import hostsim
from parity import COST_TOL, U0_TOL, VIOL_TOL, X1_TOL, golden_errors


class P:
    pass


def problem(g, k, N):
    p = P()
    p.N = N
    p.x0 = g["x0"][k]
    p.com_ref = g["com_ref"][k].T
    f = g["foot_ref"][k]
    p.pl_ref, p.pr_ref, p.al_ref, p.ar_ref = f[:, 0:3].T, f[:, 3:6].T, f[:, 6], f[:, 7]
    p.gl, p.gr = g["gamma"][k][:, 0], g["gamma"][k][:, 1]
    per = np.ndim(g["mass"]) > 0                                   # round-2 golden files carry per-instance mass / k1
    p.mass, p.k1 = (float(g["mass"][k]), float(g["k1"][k])) if per else (float(g["mass"]), float(g["k1"]))
    p.eps_reg, p.w_rate = None, 1.0                               # eps_reg: product default
    return p


@pytest.mark.parametrize("N", [10, 20])
def test_core_matches_oracle_golden(golden, N):
    g = golden[N]
    for k in range(len(g["ticks"])):
        if g["status"][k] != 0:
            continue
        r = hostsim.solve(problem(g, k, N))
        assert r["status"] == 0, (N, int(g["ticks"][k]), r["status"])
        assert r["viol"] <= VIOL_TOL
        ec, ex, eu = golden_errors(g, k, r["cost"], r["X"][:20, 1], r["U"][:, 0])
        assert ec <= COST_TOL and ex <= X1_TOL and eu <= U0_TOL, (N, int(g["ticks"][k]), ec, ex, eu)


@pytest.mark.parametrize("name,N", [("payload_N10", 10), ("perturbed_N20", 20), ("N60", 60)])
def test_core_matches_round2_golden(golden, name, N):
    """Payload variant (k1 = 7, masses 40.05 / 45 / 50), perturbed initial states (BASELINE config 3 recipe) and the long
    horizon (config 5) against oracle-T (tests/golden/make_golden_r2.py)."""
    g = golden[name]
    for k in range(len(g["ticks"])):
        r = hostsim.solve(problem(g, k, N))
        assert r["status"] == 0, (name, int(g["ticks"][k]), r["status"])
        assert r["viol"] <= VIOL_TOL
        ec, ex, eu = golden_errors(g, k, r["cost"], r["X"][:20, 1], r["U"][:, 0])
        assert ec <= COST_TOL and ex <= X1_TOL and eu <= U0_TOL, (name, int(g["ticks"][k]), ec, ex, eu)


def test_analytic_stage_hessian_against_finite_differences(golden):
    """Hessian of the stage Lagrangian assembled by the backward pass == central differences of the analytic
    stage gradient (costates and multipliers random, slacks huge so the barrier part vanishes)."""
    g = golden[10]
    N, NX, NU, NR = 10, 28, 32, 56
    k = list(g["ticks"]).index(262)                    # a landing inside the horizon: both feet, swing, yaw terms
    prob = problem(g, k, N)
    r = hostsim.solve(prob, max_iter=8)
    L = hostsim.lib()
    x0, com, foot, gam = hostsim.pack(prob)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    oU = (N + 1) * NX; oY = oU + N * NU; oS = oY + (N + 1) * NX; oL = oS + (N + 1) * NR
    rng = np.random.default_rng(1)
    wk = r["work"].copy()
    wk[oS:oL] = 1e30
    wk[oY:oS] = rng.normal(size=(N + 1) * NX) * 50

    def dbg(work, i):
        gr, M = np.zeros(60), np.zeros(3600)
        L.hostsim_stage_debug(ctypes.c_int(N), dp(x0), dp(com), dp(foot), dp(gam), ctypes.c_double(prob.mass),
                              ctypes.c_double(prob.k1), dp(work), ctypes.c_int(i), ctypes.c_double(1e-3), dp(gr), dp(M))
        return gr, M.reshape(60, 60)

    for i in (1, 4, 8):
        _, M = dbg(wk, i)
        H = np.zeros((60, 60))
        for j in range(60):
            h = 1e-6
            wp, wm = wk.copy(), wk.copy()
            idx = oU + i * NU + j if j < 32 else i * NX + (j - 32)
            wp[idx] += h; wm[idx] -= h
            H[:, j] = (dbg(wp, i)[0] - dbg(wm, i)[0]) / (2 * h)
        assert np.abs(H - M).max() <= 1e-5 * max(1.0, np.abs(M).max()), i


def test_wide_eval_pass_matches_thread_per_stage_evaluation(golden):
    """The CTA-wide eval and trial passes (stage x role work items) give the same derivative record, statistics and merit
    quantities as the plain thread-per-stage evaluations they replaced, at mid-solve iterates of standing, swing and
    landing ticks."""
    for N in (10, 20):
        g = golden[N]
        for tick in (0, 262, 230):
            if tick not in list(g["ticks"]):
                continue
            k = list(g["ticks"]).index(tick)
            prob = problem(g, k, N)
            r = hostsim.solve(prob, max_iter=6)
            L = hostsim.lib()
            x0, com, foot, gam = hostsim.pack(prob)
            dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            out = np.zeros(9)
            L.hostsim_eval_compare(ctypes.c_int(N), dp(x0), dp(com), dp(foot), dp(gam), ctypes.c_double(prob.mass),
                                   ctypes.c_double(prob.k1), dp(r["work"]), dp(out))
            assert out[0] <= 1e-10 and out[1] <= 1e-9 and out[8] <= 1e-10, (N, tick, out)     # summation order differs


def test_crawling_warm_start_is_abandoned():
    """Tick 867 of the recorded N = 20 walk (inside the push window): the warm attempt takes ~50 steps of length 1e-2 .. 1e-6 before it
    gets going (74 iterations in all; a cold start needs 36).  The crawl rule (`crawl_window` consecutive steps shorter than
    `crawl_alpha`) restarts it cold: the solve ends at the same KKT point in far fewer iterations; with the rule off the old count
    comes back."""
    w = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "walk_ticks_N20.npz")))
    t = 867
    r2 = hostsim.solve(problem(w, t - 2, 20))
    r1 = hostsim.solve(problem(w, t - 1, 20), work=r2["work"].copy(), warm=4)
    on = hostsim.solve(problem(w, t, 20), work=r1["work"].copy(), warm=4)
    off = hostsim.solve(problem(w, t, 20), work=r1["work"].copy(), warm=4, crawl_window=0)
    cold = hostsim.solve(problem(w, t, 20))
    assert on["status"] == 0 and off["status"] == 0 and cold["status"] == 0
    assert off["iters"] >= 65 and on["iters"] <= 58 and cold["iters"] <= 40, (on["iters"], off["iters"], cold["iters"])
    assert abs(on["cost"] - off["cost"]) <= COST_TOL * max(1.0, abs(off["cost"]))
    assert abs(on["cost"] - cold["cost"]) <= COST_TOL * max(1.0, abs(cold["cost"]))


def test_stalled_warm_start_at_the_short_horizon_is_abandoned_early():
    """Tick 1574 of the recorded N = 10 walk: the warm attempt sits at step lengths of 0.7 without halving its error (60 iterations under
    the general stall window, 90 in all; a cold start needs 30).  At N <= 12 a warm attempt gets a window of 25 (`warm_stall_window`):
    same KKT point, 55 iterations."""
    w = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "walk_ticks_N10.npz")))
    t = 1574
    r2 = hostsim.solve(problem(w, t - 2, 10))
    r1 = hostsim.solve(problem(w, t - 1, 10), work=r2["work"].copy(), warm=4)
    on = hostsim.solve(problem(w, t, 10), work=r1["work"].copy(), warm=4)
    off = hostsim.solve(problem(w, t, 10), work=r1["work"].copy(), warm=4, warm_stall_window=0)
    assert on["status"] == 0 and off["status"] == 0
    assert off["iters"] >= 80 and on["iters"] <= 60, (on["iters"], off["iters"])
    assert abs(on["cost"] - off["cost"]) <= COST_TOL * max(1.0, abs(off["cost"]))
