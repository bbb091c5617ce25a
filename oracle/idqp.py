"""ORACLE (test infrastructure) -- the whole-body inverse-dynamics QP of the reference, restated in numpy.

`build_qp` follows `InverseDynamics.get_joint_torques` (`code/inverse_dynamics.py:30-135`) line by line from the point
where DART has delivered the Jacobians (:46-66), errors (:77-92), inertia matrix and bias forces (:113-118): cost
blocks (:95-108), contact-force regularisation (:111), equations of motion (:113-118), centre-of-pressure / friction
rows (:120-130).  `solve_qp` replaces `QPSolver.solve` (`code/utils.py:40-92`, CasADi conic + OSQP, not installable here)
by a plain dense primal-dual interior-point method on the full KKT system with slacks (numpy.linalg.solve), run to 1e-11;
OSQP itself stops at its default eps 1e-3, so the parity target is the exact optimum, as for the MPC.

Parity status: UNPINNED against OSQP (no fixture in the reference tree); cross-checked against scipy SLSQP in
tests/test_idqp.py.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

TASKS = ["lfoot", "rfoot", "com", "torso", "base", "joints"]                                        # :41
WEIGHTS = {"lfoot": 1., "rfoot": 1., "com": 1., "torso": 1., "base": 1., "joints": 1.e-1}            # :42
POS_GAINS = {"lfoot": 10., "rfoot": 10., "com": 5., "torso": 10., "base": 10., "joints": 10.}        # :43
VEL_GAINS = {"lfoot": 5., "rfoot": 5., "com": 10., "torso": 5., "base": 3., "joints": 5}             # :44


def block_diag(*ms):
    r = sum(m.shape[0] for m in ms); c = sum(m.shape[1] for m in ms)
    out = np.zeros((r, c)); i = j = 0
    for m in ms:
        out[i:i + m.shape[0], j:j + m.shape[1]] = m; i += m.shape[0]; j += m.shape[1]
    return out


def build_qp(dofs, J, Jdot, ff, pos_error, vel_error, qdot, inertia, bias, contact, foot_size=0.1, mu=0.5):
    """H, F, A_eq, b_eq, A_ineq, b_ineq exactly as handed to `QPSolver.set_values` (:133)."""
    contact_l = contact in ("lfoot", "ds")                                                          # :31-32
    contact_r = contact in ("rfoot", "ds")
    d = foot_size / 2.0                                                                             # :9
    n_vars = 2 * dofs + 12                                                                          # :13-15
    H = np.zeros((n_vars, n_vars)); F = np.zeros(n_vars)
    qdd = np.arange(dofs); fc = np.arange(2 * dofs, n_vars)
    for task in TASKS:                                                                              # :101-108
        H_task = WEIGHTS[task] * J[task].T @ J[task]
        F_task = -WEIGHTS[task] * J[task].T @ (ff[task] + VEL_GAINS[task] * vel_error[task] + POS_GAINS[task] * pos_error[task]
                                               - Jdot[task] @ qdot)
        H[np.ix_(qdd, qdd)] += H_task
        F[qdd] += F_task
    H[np.ix_(fc, fc)] += np.eye(len(fc)) * 1e-6                                                     # :111
    actuation = block_diag(np.zeros((6, 6)), np.eye(dofs - 6))                                      # :115
    contact_jac = np.vstack((contact_l * J["lfoot"], contact_r * J["rfoot"]))                       # :116
    A_eq = np.hstack((inertia, -actuation, -contact_jac.T))                                         # :117
    b_eq = -bias                                                                                    # :118
    A = np.array([[1, 0, 0, 0, 0, -d], [-1, 0, 0, 0, 0, -d], [0, 1, 0, 0, 0, -d], [0, -1, 0, 0, 0, -d],
                  [0, 0, 0, 1, 0, -mu], [0, 0, 0, -1, 0, -mu], [0, 0, 0, 0, 1, -mu], [0, 0, 0, 0, -1, -mu]], float)   # :123-130
    A_in = np.zeros((16, n_vars)); b_in = np.zeros(16)
    A_in[:, fc] = block_diag(A, A)                                                                  # :131
    return H, F, A_eq, b_eq, A_in, b_in


def solve_qp(H, F, A_eq, b_eq, A_in, b_in, tol=1e-11, eps=1e-9, max_iter=200):
    """min 1/2 x'Hx + F'x, A_eq x = b_eq, A_in x <= b_in.  Full (unreduced) KKT system with slacks, fixed centring 0.1."""
    n, me, mi = len(F), len(b_eq), len(b_in)
    x, y, z, s = np.zeros(n), np.zeros(me), np.ones(mi), np.maximum(b_in, 1.0)
    Hr = H + eps * np.eye(n)
    for it in range(max_iter):
        rd = Hr @ x + F + A_eq.T @ y + A_in.T @ z
        rp = A_eq @ x - b_eq
        ri = A_in @ x + s - b_in
        mu = s @ z / max(mi, 1)
        if max(np.abs(rd).max(), np.abs(rp).max(initial=0), np.abs(ri).max(initial=0)) <= tol * max(1.0, np.abs(F).max(), np.abs(b_eq).max(initial=0)) and mu <= tol:
            return x, 0, it
        K = np.zeros((n + me + 2 * mi,) * 2)
        K[:n, :n] = Hr; K[:n, n:n + me] = A_eq.T; K[:n, n + me:n + me + mi] = A_in.T
        K[n:n + me, :n] = A_eq
        K[n + me:n + me + mi, :n] = A_in; K[n + me:n + me + mi, n + me + mi:] = np.eye(mi)
        K[n + me + mi:, n + me:n + me + mi] = np.diag(s); K[n + me + mi:, n + me + mi:] = np.diag(z)
        rhs = -np.concatenate([rd, rp, ri, s * z - 0.1 * mu])
        dlt = np.linalg.solve(K, rhs)
        dx, dy, dz, ds = dlt[:n], dlt[n:n + me], dlt[n + me:n + me + mi], dlt[n + me + mi:]
        a = 1.0
        for v, dv in ((s, ds), (z, dz)):
            neg = dv < 0
            if neg.any():
                a = min(a, 0.99 * float(np.min(-v[neg] / dv[neg])))
        x, y, z, s = x + a * dx, y + a * dy, z + a * dz, s + a * ds
    return x, 1, max_iter


def synthetic_task(rng, dofs=30, contact="ds", scale=1.0):
    """Random but physically shaped data in place of the DART calls (:46-66, :113-118): full-rank Jacobians, SPD inertia."""
    J = {"lfoot": rng.normal(size=(6, dofs)), "rfoot": rng.normal(size=(6, dofs)), "com": rng.normal(size=(3, dofs)),
         "torso": rng.normal(size=(3, dofs)), "base": rng.normal(size=(3, dofs))}
    sel = np.zeros((dofs, dofs))
    for i in rng.choice(np.arange(6, dofs), size=min(12, dofs - 6), replace=False):                  # redundant dofs (:23-28)
        sel[i, i] = 1.0
    J["joints"] = sel
    Jdot = {k: 0.1 * rng.normal(size=v.shape) for k, v in J.items()}
    Jdot["joints"] = np.zeros((dofs, dofs))                                                        # :64
    ff = {k: scale * rng.normal(size=v.shape[0]) for k, v in J.items()}
    pe = {k: 0.05 * scale * rng.normal(size=v.shape[0]) for k, v in J.items()}
    ve = {k: 0.1 * scale * rng.normal(size=v.shape[0]) for k, v in J.items()}
    A = rng.normal(size=(dofs, dofs))
    inertia = A @ A.T / dofs + np.diag(rng.uniform(0.5, 2.0, dofs))
    bias = rng.normal(size=dofs) * 5.0
    bias[2] += 40.0 * 9.81                                                                          # weight on the floating base z
    qdot = 0.2 * rng.normal(size=dofs)
    return dict(dofs=dofs, J=J, Jdot=Jdot, ff=ff, pos_error=pe, vel_error=ve, qdot=qdot, inertia=inertia, bias=bias, contact=contact)
