"""Whole-body inverse-dynamics QP: the numpy restatement of `inverse_dynamics.py:95-131` against the
vectorised product builder and an independent solver (CPU), and the CUDA QP kernel against the oracle (GPU)."""
import numpy as np
import pytest


def _batch(rng, B, dofs=30):
    from oracle.idqp import synthetic_task
    contacts = ["ds", "lfoot", "rfoot"]
    return [synthetic_task(rng, dofs=dofs, contact=contacts[b % 3]) for b in range(B)]


def _stack(tasks):
    keys = tasks[0]["J"].keys()
    st = lambda name: {k: np.stack([t[name][k] for t in tasks]) for k in keys}
    return dict(J=st("J"), Jdot=st("Jdot"), ff=st("ff"), pos_error=st("pos_error"), vel_error=st("vel_error"),
                qdot=np.stack([t["qdot"] for t in tasks]), inertia=np.stack([t["inertia"] for t in tasks]),
                bias=np.stack([t["bias"] for t in tasks]))


def test_batched_builder_matches_the_literal_restatement(pkg):
    from oracle.idqp import build_qp
    from cmpc_b200.idqp import assemble_id_qp
    rng = np.random.default_rng(0)
    tasks = _batch(rng, 6)
    s = _stack(tasks)
    cl = np.array([t["contact"] in ("lfoot", "ds") for t in tasks]); cr = np.array([t["contact"] in ("rfoot", "ds") for t in tasks])
    mats = assemble_id_qp(s["J"], s["Jdot"], s["ff"], s["pos_error"], s["vel_error"], s["qdot"], s["inertia"], s["bias"], cl, cr)
    for b, t in enumerate(tasks):
        ref = build_qp(**t)
        for a, r in zip(mats, ref):
            assert np.abs(a[b] - r).max() <= 1e-12 * max(1.0, np.abs(r).max())


def test_oracle_qp_against_scipy(pkg):
    """The oracle's interior point and scipy's SLSQP agree on the actuated torques and the cost (two unrelated methods)."""
    from scipy.optimize import minimize
    from oracle.idqp import build_qp, solve_qp
    rng = np.random.default_rng(1)
    for t in _batch(rng, 3, dofs=12):                                 # small dofs: SLSQP is a dense method
        H, F, Ae, be, Ai, bi = build_qp(**t)
        x, st, it = solve_qp(H, F, Ae, be, Ai, bi)
        assert st == 0
        Hr = H + 1e-9 * np.eye(len(F))
        res = minimize(lambda v: 0.5 * v @ Hr @ v + F @ v, np.zeros(len(F)), jac=lambda v: Hr @ v + F, method="SLSQP",
                       constraints=[dict(type="eq", fun=lambda v: Ae @ v - be, jac=lambda v: Ae), dict(type="ineq", fun=lambda v: bi - Ai @ v, jac=lambda v: -Ai)],
                       options=dict(ftol=1e-14, maxiter=500))
        d = t["dofs"]
        cost = lambda v: 0.5 * v @ H @ v + F @ v
        assert abs(cost(res.x) - cost(x)) <= 1e-6 * max(1.0, abs(cost(x)))
        assert np.abs(res.x[d + 6:2 * d] - x[d + 6:2 * d]).max() <= 1e-4 * max(1.0, np.abs(x[d + 6:2 * d]).max())
        assert np.abs(Ae @ x - be).max() <= 1e-8 and (Ai @ x - bi).max() <= 1e-8


@pytest.mark.gpu
def test_gpu_qp_matches_the_oracle(pkg):
    from oracle.idqp import build_qp, solve_qp
    from cmpc_b200.idqp import QPSolver, joint_torques
    rng = np.random.default_rng(2)
    B, dofs = 48, 30                                                  # HRP-4: 24 joints + 6 floating-base coordinates
    tasks = _batch(rng, B, dofs)
    mats = [build_qp(**t) for t in tasks]
    qp = QPSolver(2 * dofs + 12, dofs, 16, batch=B)
    qp.set_values(*[np.stack([m[k] for m in mats]) for k in range(6)])
    x = qp.solve()
    assert (qp.status == 0).all(), qp.status
    for b, (H, F, Ae, be, Ai, bi) in enumerate(mats):
        xo, st, _ = solve_qp(H, F, Ae, be, Ai, bi)
        assert st == 0
        cost = lambda v: 0.5 * v @ H @ v + F @ v
        assert abs(cost(x[b]) - cost(xo)) <= 1e-7 * max(1.0, abs(cost(xo))), b
        tau, tau_o = x[b, dofs + 6:2 * dofs], xo[dofs + 6:2 * dofs]
        assert np.abs(tau - tau_o).max() <= 1e-6 * max(1.0, np.abs(tau_o).max()), b      # the returned quantity (:135-136)
        assert np.abs(Ae @ x[b] - be).max() <= 1e-7 and (Ai @ x[b] - bi).max() <= 1e-7
    s = _stack(tasks)
    tau, status = joint_torques(s["J"], s["Jdot"], s["ff"], s["pos_error"], s["vel_error"], s["qdot"], s["inertia"], s["bias"], [t["contact"] for t in tasks])
    assert (status == 0).all() and np.abs(tau - x[:, dofs + 6:2 * dofs]).max() <= 1e-9
    # one QP through the single-instance surface of the reference class
    one = QPSolver(2 * dofs + 12, dofs, 16)
    one.set_values(*mats[0])
    assert np.abs(one.solve() - x[0]).max() <= 1e-12
