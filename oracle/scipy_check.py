"""ORACLE (test infrastructure) -- independent cross-check of oracle-T with scipy's NLP solvers.

The LITERAL NLP of `code/centroidal_mpc_vertices.py:185-353` (oracle/spec.py: symbols -> sympy derivatives) is handed to
`scipy.optimize.minimize` -- `trust-constr` (Byrd-Hribar-Nocedal interior point / trust region, exact Hessians) or
`SLSQP` (Kraft's sequential least-squares QP) -- started from oracle-T's answer perturbed by 1e-3 relative.  Neither
solver shares any code with the builder's interior-point methods (oracle/ipm_py.py, oracle/ipm_c.c, the CUDA product):
if they return to oracle-T's point, that point is a KKT point of the literal NLP by two unrelated algorithms.

trust-constr stalls on the standing ticks (LICQ fails there: Lyapunov rows with zero value and zero gradient, SURVEY.md
section 7); those instances are checked with SLSQP, which converges on them.

Only tests/ and tests/golden/ scripts may import this module.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sps
from scipy.optimize import NonlinearConstraint, minimize

from . import ipm_py

NS = 52
RELAX = 1e-8          # IPOPT bound_relax_factor, as in oracle-T


def stack_w(N, X, U):
    w = np.zeros(NS * N + 20)
    for i in range(N + 1):
        w[NS * i:NS * i + 20] = X[i]
    for i in range(N):
        w[NS * i + 20:NS * i + 52] = U[i]
    return w


def cross_check(N, x0, com_ref, foot_ref, gamma, mass, k1, X_star, U_star, method="trust-constr", perturb=1e-3, seed=0,
                maxiter=None, u0_metric=None):
    """Returns a dict with the distances between scipy's answer and (X_star, U_star) in the parity metrics."""
    prob = ipm_py.unpack_problem(N, x0, com_ref, foot_ref, gamma, mass, k1)
    nlp = ipm_py.NLP(prob)
    w_star = stack_w(N, X_star, U_star)
    rng = np.random.default_rng(seed)
    w0 = w_star + perturb * rng.standard_normal(nlp.n) * np.maximum(1.0, np.abs(w_star))
    me, mi = nlp.m_eq, nlp.m_in
    t0 = time.time()
    if method == "trust-constr":
        Hf = lambda w: nlp.hess(w, np.zeros(me), np.zeros(mi))
        con = NonlinearConstraint(lambda w: np.concatenate([nlp.eq(w), nlp.ineq(w)]),
                                  np.concatenate([np.zeros(me), np.full(mi, -np.inf)]), np.concatenate([np.zeros(me), np.full(mi, RELAX)]),
                                  jac=lambda w: sps.vstack([nlp.jac_eq(w), nlp.jac_in(w)]).tocsr(),
                                  hess=lambda w, v: nlp.hess(w, v[:me], v[me:]) - Hf(w))
        res = minimize(nlp.cost, w0, jac=nlp.grad, hess=Hf, constraints=[con], method="trust-constr",
                       options=dict(gtol=1e-9, xtol=1e-12, barrier_tol=1e-9, maxiter=maxiter or 3000, initial_barrier_parameter=1e-3,
                                    initial_barrier_tolerance=1e-3))
    elif method == "SLSQP":
        cons = [dict(type="eq", fun=nlp.eq, jac=lambda w: nlp.jac_eq(w).toarray()),
                dict(type="ineq", fun=lambda w: RELAX - nlp.ineq(w), jac=lambda w: -nlp.jac_in(w).toarray())]
        res = minimize(nlp.cost, w0, jac=nlp.grad, constraints=cons, method="SLSQP", options=dict(ftol=1e-14, maxiter=maxiter or 500))
    else:
        raise ValueError(method)
    w = res.x
    Js, J = nlp.cost_reference(w_star), nlp.cost_reference(w)
    if u0_metric is None:                                      # plain relative 2-norm of the first applied contact forces
        u0_metric = lambda u, us, x0_, g0: [np.linalg.norm(u[:24] - us[:24]) / max(np.linalg.norm(us[:24]), 1e-9)]
    return {"method": method, "nit": int(res.nit), "message": str(res.message), "seconds": time.time() - t0,
            "start_dist": float(np.abs(w0 - w_star).max()), "cost": float(J), "cost_star": float(Js),
            "cost_err": float(abs(J - Js) / max(1.0, abs(Js))), "x1_err": float(np.abs(w[NS:NS + 12] - w_star[NS:NS + 12]).max()),
            "u0_err": float(u0_metric(w[20:52], w_star[20:52], np.asarray(x0), np.asarray(gamma).reshape(N + 1, 2)[0])[0]),
            "viol": float(nlp.violation(w)), "viol_star": float(nlp.violation(w_star))}
