"""Parity metrics shared by the tests, smoke() and bench.py (definitions in DESIGN.md section 'Parity').

  cost   |J - J*| <= 1e-6 * max(1, |J*|)                 (J* ~ 1e-9 at standing ticks -> absolute floor)
  u0     first applied vertex forces u0[0:24], relative 2-norm <= 1e-4, compared MODULO the internal force
         between the feet: in double support an equal and opposite force on the two feet along the line
         joining the foot centres produces no net wrench, costs nothing in the reference cost (:320-351 penalise
         only deviations from the per-foot mean and f_z rates) and is therefore not determined by the NLP.
         u0[24:32] (foot velocities / yaw rates) carry no cost and, for a stance foot, no effect: not compared.
  viol   max unrelaxed violation of all rows and dynamics defects <= 1e-6
"""
import numpy as np

COST_TOL, U0_TOL, VIOL_TOL, X1_TOL = 1e-6, 1e-4, 1e-6, 1e-6


def cost_err(J, Jstar):
    return np.abs(J - Jstar) / np.maximum(1.0, np.abs(Jstar))


def u0_err(u0, u0_star, x0, gamma0):
    """relative error of u0[0:24] after projecting out the wrench-free internal force (double support only)."""
    u0 = np.atleast_2d(u0)[:, :24]
    u0s = np.atleast_2d(u0_star)[:, :24]
    x0 = np.atleast_2d(x0)
    gamma0 = np.atleast_2d(gamma0)
    d = u0 - u0s
    e = x0[:, 13:16] - x0[:, 17:20]
    e = e / np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-12)
    n = np.concatenate([np.tile(e, (1, 4)), np.tile(-e, (1, 4))], axis=1) / np.sqrt(8.0)
    ds = (gamma0[:, 0] > 0.5) & (gamma0[:, 1] > 0.5)
    d = np.where(ds[:, None], d - n * np.sum(n * d, axis=1, keepdims=True), d)
    # forces of a swing foot are pinned to ~0 by the 10|f|^2 term; they are part of the comparison as they are
    return np.linalg.norm(d, axis=1) / np.maximum(np.linalg.norm(u0s, axis=1), 1e-9)


def golden_errors(g, k, cost, x1, u0):
    """Errors of one result against golden instance k: the best match over the oracle's KKT points of the instance
    (the NLP is non-convex; tests/golden/make_golden_alts.py stores the alternatives)."""
    if "cost_alt" in g:
        cands = [(g["cost_alt"][k, a], g["X_alt"][k, a], g["U_alt"][k, a]) for a in range(g["cost_alt"].shape[1])
                 if np.isfinite(g["cost_alt"][k, a])]
    else:
        cands = [(g["cost"][k], g["X"][k], g["U"][k])]
    best = None
    for c, X, U in cands:
        e = (float(cost_err(cost, c)), float(np.abs(x1[:12] - X[1, :12]).max()),
             float(u0_err(u0, U[0], g["x0"][k], g["gamma"][k, 0])[0]))
        score = max(e[0] / COST_TOL, e[1] / X1_TOL, e[2] / U0_TOL)
        if best is None or score < best[0]:
            best = (score,) + e
    return best[1:]
