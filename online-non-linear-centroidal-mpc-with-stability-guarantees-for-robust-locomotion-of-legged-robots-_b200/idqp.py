"""Batched whole-body inverse-dynamics QP on the GPU: the step of the reference's control loop that
consumes the MPC's CoM acceleration (`simulation.py:208-210, :276`).

  reference                                                  here
  utils.QPSolver(n_vars, n_eq, n_ineq)      (utils.py:40-73) QPSolver(n_vars, n_eq, n_ineq, batch)
  .set_values(H, F, A_eq, b_eq, A_in, b_in) (:75-84)         same, arrays with a leading batch axis (or none for one QP)
  .solve() -> x, zeros on failure           (:85-92)         same; per-QP status in `.status` (cmpc_qp_solve_host)
  InverseDynamics.get_joint_torques         (inverse_dynamics.py:30-135)
                                                             `assemble_id_qp` builds the six matrices of :95-131 for a batch of robots
                                                             from the Jacobians / errors / inertia a rigid-body library delivers
                                                             (DART in the reference, :46-66, :113-118), `joint_torques` returns tau[6:]

The solve is the CUDA kernel behind the C ABI (csrc/cmpc_qp.cuh); there is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

TASKS = ("lfoot", "rfoot", "com", "torso", "base", "joints")                                         # inverse_dynamics.py:41-44
WEIGHTS = {"lfoot": 1., "rfoot": 1., "com": 1., "torso": 1., "base": 1., "joints": 1.e-1}
POS_GAINS = {"lfoot": 10., "rfoot": 10., "com": 5., "torso": 10., "base": 10., "joints": 10.}
VEL_GAINS = {"lfoot": 5., "rfoot": 5., "com": 10., "torso": 5., "base": 3., "joints": 5.}


class QPSolver:
    def __init__(self, n_vars, n_eq_constraints=0, n_ineq_constraints=0, batch=1, device=0, tol=1e-9, max_iter=60):
        self.n_vars, self.n_eq_constraints, self.n_ineq_constraints = int(n_vars), int(n_eq_constraints), int(n_ineq_constraints)
        self.batch, self.device, self.tol, self.max_iter = int(batch), int(device), float(tol), int(max_iter)
        self._L = _lib.load()
        self._L.cmpc_qp_solve_host.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 6 + [ctypes.c_double, ctypes.c_int32] + [ctypes.c_void_p] * 3
        self._v = None
        self.status = None
        self.iters = None

    def set_values(self, H, F, A_eq=None, b_eq=None, A_ineq=None, b_ineq=None):
        B, n, me, mi = self.batch, self.n_vars, self.n_eq_constraints, self.n_ineq_constraints
        arr = lambda a, shp: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float64).reshape((-1,) + shp), (B,) + shp))
        self._v = (arr(H, (n, n)), arr(F, (n,)),
                   arr(A_eq, (me, n)) if me else np.zeros((B, 0, n)), arr(b_eq, (me,)) if me else np.zeros((B, 0)),
                   arr(A_ineq, (mi, n)) if mi else np.zeros((B, 0, n)), arr(b_ineq, (mi,)) if mi else np.zeros((B, 0)))

    def solve(self):
        if self._v is None:
            raise _lib.CmpcError("QPSolver.solve before set_values")
        B, n = self.batch, self.n_vars
        x = np.empty((B, n)); st = np.empty(B, np.int32); it = np.empty(B, np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = self._L.cmpc_qp_solve_host(self.device, B, n, self.n_eq_constraints, self.n_ineq_constraints, *[p(a) for a in self._v],
                                        self.tol, self.max_iter, p(x), p(st), p(it))
        _lib._check(self._L, rc, "cmpc_qp_solve_host")
        self.status, self.iters = st, it
        x[st != 0] = 0.0                                           # the reference returns zeros when its solver fails (utils.py:85-92)
        return x if B > 1 else x[0]


def assemble_id_qp(J, Jdot, ff, pos_error, vel_error, qdot, inertia, bias, contact_l, contact_r, foot_size=0.1, mu=0.5):
    """The six QP matrices of `get_joint_torques` (:95-131) for a batch of B robots.
    J / Jdot: dicts task -> [B, m_task, dofs]; ff / pos_error / vel_error: dicts task -> [B, m_task]; qdot [B, dofs];
    inertia [B, dofs, dofs]; bias = Coriolis + gravity forces [B, dofs]; contact_l / contact_r: bool [B]."""
    B, dofs = qdot.shape
    n = 2 * dofs + 12
    H = np.zeros((B, n, n)); F = np.zeros((B, n))
    for task in TASKS:                                                                              # :101-108
        Jt = J[task]
        acc = ff[task] + VEL_GAINS[task] * vel_error[task] + POS_GAINS[task] * pos_error[task] - np.einsum("bmd,bd->bm", Jdot[task], qdot)
        H[:, :dofs, :dofs] += WEIGHTS[task] * np.einsum("bmi,bmj->bij", Jt, Jt)
        F[:, :dofs] -= WEIGHTS[task] * np.einsum("bmi,bm->bi", Jt, acc)
    idx = np.arange(2 * dofs, n)
    H[:, idx, idx] += 1e-6                                                                          # :111
    cl = np.asarray(contact_l, float)[:, None, None]; cr = np.asarray(contact_r, float)[:, None, None]
    A_eq = np.zeros((B, dofs, n))
    A_eq[:, :, :dofs] = inertia
    A_eq[:, np.arange(6, dofs), dofs + np.arange(6, dofs)] = -1.0                                   # :115  -block_diag(0_6, I)
    A_eq[:, :, 2 * dofs:2 * dofs + 6] = -np.transpose(cl * J["lfoot"], (0, 2, 1))                   # :116-117
    A_eq[:, :, 2 * dofs + 6:] = -np.transpose(cr * J["rfoot"], (0, 2, 1))
    b_eq = -np.asarray(bias, float)                                                                 # :118
    d = foot_size / 2.0
    A = np.array([[1, 0, 0, 0, 0, -d], [-1, 0, 0, 0, 0, -d], [0, 1, 0, 0, 0, -d], [0, -1, 0, 0, 0, -d],
                  [0, 0, 0, 1, 0, -mu], [0, 0, 0, -1, 0, -mu], [0, 0, 0, 0, 1, -mu], [0, 0, 0, 0, -1, -mu]], float)   # :123-130
    A_in = np.zeros((B, 16, n))
    A_in[:, 0:8, 2 * dofs:2 * dofs + 6] = A
    A_in[:, 8:16, 2 * dofs + 6:] = A                                                                # :131
    return H, F, A_eq, b_eq, A_in, np.zeros((B, 16))


def joint_torques(J, Jdot, ff, pos_error, vel_error, qdot, inertia, bias, contact, device=0, **kw):
    """Batched `get_joint_torques`: contact is a list of 'ds' | 'lfoot' | 'rfoot' per robot.  Returns (tau[6:] [B, dofs-6], status)."""
    cl = np.array([c in ("lfoot", "ds") for c in contact]); cr = np.array([c in ("rfoot", "ds") for c in contact])      # :31-32
    mats = assemble_id_qp(J, Jdot, ff, pos_error, vel_error, qdot, inertia, bias, cl, cr, **kw)
    B, dofs = qdot.shape
    qp = QPSolver(2 * dofs + 12, dofs, 16, batch=B, device=device)
    qp.set_values(*mats)
    x = qp.solve().reshape(B, -1)
    return x[:, dofs + 6:2 * dofs], qp.status                                                       # :135-136
