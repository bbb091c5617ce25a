"""Drop-in for the reference module `code/centroidal_mpc_vertices.py`: same class name, constructor and
`solve(current, t)` contract (:7, :480-683), with the CasADi Opti build/solve replaced by the B200 solver.

  reference                                   here
  cs.Opti() + NLP build (:126-353)            BatchSolver(N, 1) -> cmpc_create
  opt.set_value(...) x (4N + 4) (:511-600)    assembly.assemble_tick (table look-ups)
  opt.solve() (:606)                          cmpc_solve_host_traj (H2D, solve, D2H of x1 / u0 and the trajectories:
  sol.value(...) (:614-619)                   one call and one synchronisation per control tick)
  opt.set_initial(...) (:630-631)             warm-start state stays on the device inside the handle
"""
from __future__ import annotations

import numpy as np

from ._lib import COLD, WARM_AUTO, WARM_FULL, WARM_PRIMAL, BatchSolver, STATUS_NAMES  # noqa: F401
from .assembly import PlanTables, ReferenceTables, assemble_tick

_K1_DEFAULT = (4.0, 0.1)      # (:27-28)
_K1_RATE10 = (5.0, 0.2)       # (:29-31)


class centroidal_mpc:  # noqa: N801  (reference class name)
    K1K2 = None               # the payload module overrides this with (7, 1)

    def __init__(self, initial, footstep_planner, params, CoM_ref, contact_trj_l=None, contact_trj_r=None,
                 device=0, warm_mode=WARM_AUTO, **solver_overrides):
        self.params = params
        self.N = params["N"]
        self.delta = params["world_time_step"] * params["mpc_rate"]
        self.h = params.get("h")
        self.eta = params.get("eta")
        self.foot_size = params.get("foot_size")
        self.mass = params["mass"]
        self.g = params["g"]
        self.initial = initial
        self.footstep_planner = footstep_planner
        self.debug = 0
        self.update_contact_flag = 0
        if self.K1K2 is not None:
            self.k1, self.k2 = self.K1K2
        else:
            self.k1, self.k2 = _K1_RATE10 if params["mpc_rate"] == 10 else _K1_DEFAULT
        self.mpc_rate = params["mpc_rate"]
        self.update_swing_trj = 0
        self.contact_trj_l, self.contact_trj_r = contact_trj_l, contact_trj_r
        self._tables = PlanTables(footstep_planner.plan)
        self._refs = ReferenceTables(CoM_ref, footstep_planner)
        self._warm_mode = warm_mode
        self._first = True
        w_rate = 0.0 if self.mpc_rate == 10 else 1.0               # :339-341
        self._solver = BatchSolver(self.N, 1, device=device, delta=self.delta, grav=self.g, w_rate=w_rate,
                                   **solver_overrides)
        self.current_state = np.zeros(20)
        self.model_state = {"com": {"pos": np.zeros(3), "vel": np.zeros(3), "acc": np.zeros(3)},
                            "hw": {"val": np.zeros(3), "dot": np.zeros(3)},
                            "theta_hat": {"val": np.zeros(3)},
                            "ang_contact_left": {"val": np.zeros(3)}, "pos_contact_left": {"val": np.zeros(3)},
                            "ang_contact_right": {"val": np.zeros(3)}, "pos_contact_right": {"val": np.zeros(3)},
                            "mpc_new_contact": {"val": np.zeros(3)}, "counter": {"val": 0}}     # :358-366
        self.last_status = None
        self.last_iters = 0
        self.last_cost = 0.0

    def solve(self, current, t):
        fp, N = self.footstep_planner, self.N
        x0, com, foot, gamma = assemble_tick(self._tables, self._refs, fp.plan, self.params, current,
                                             self.model_state["theta_hat"]["val"], t)
        self.current_state = x0
        mode = COLD if self._first else self._warm_mode
        out = self._solver.solve_host(x0[None], com[None], foot[None], gamma[None], self.mass, self.k1, mode, traj_batch=1)
        self._first = False
        self.last_status, self.last_iters, self.last_cost = int(out["status"][0]), int(out["iters"][0]), float(out["cost"][0])
        if self.last_status != 0:                                  # the reference dies here (:605-614)
            raise RuntimeError("centroidal MPC solve failed at t=%d: %s" % (t, STATUS_NAMES.get(self.last_status)))
        X, U = out["X"], out["U"]                               # same device-to-host synchronisation as x1 / u0
        self.x = out["x1"][0].copy()                               # :614
        self.u = out["u0"][0].copy()                               # :616
        self.x_collect = X[0].T.copy()                             # :617  (20, N+1)
        self.u_collect = U[0].T.copy()
        gl0, gr0 = gamma[0]
        Vl = self.u[0:3] + self.u[3:6] + self.u[6:9] + self.u[9:12]
        Vr = self.u[12:15] + self.u[15:18] + self.u[18:21] + self.u[21:24]
        com_acc = (gl0 * Vl + gr0 * Vr) / self.mass + np.array([0.0, 0.0, -self.g])      # :636
        hdot0 = (self.x[6:9] - x0[6:9]) / self.delta               # f(x0,u0)[6:9] (:283, :619)
        ms = self.model_state
        ms["com"]["pos"] = self.x[0:3].copy()
        ms["com"]["vel"] = self.x[3:6].copy()
        ms["com"]["acc"] = com_acc
        ms["hw"]["val"] = self.x[6:9].copy()
        ms["hw"]["dot"] = 0.01 * hdot0 * self.delta * self.mpc_rate                      # :643
        ms["theta_hat"]["val"] = self.x[9:12].copy()
        ms["ang_contact_left"]["val"] = self.x[12]
        ms["pos_contact_left"]["val"] = self.x[13:16].copy()
        ms["ang_contact_right"]["val"] = self.x[16]
        ms["pos_contact_right"]["val"] = self.x[17:20].copy()
        ms["counter"]["val"] = 0
        tb = self._tables
        if self.params["update_contact"] == "YES":                 # :656-675
            now = tb.phase_at(t)
            after = tb.phase_at(t + N * self.mpc_rate - 1)
            if now == "ss" and after == "ds" and self.update_contact_flag == 0:
                self.update_contact_flag = 1
                ms["counter"]["val"] = self.update_contact_flag
                idx = tb.step_index_at(t)
                sel = slice(17, 20) if fp.plan[idx]["foot_id"] == "lfoot" else slice(13, 16)
                fp.plan[idx + 1]["pos"] = self.x_collect[sel, N].copy()
                ms["mpc_new_contact"]["val"] = self.x_collect[sel, N].copy()
            if now == "ds":
                self.update_contact_flag = 0
        contact = tb.phase_at(t)                                    # :679-681
        if contact == "ss":
            contact = fp.plan[tb.step_index_at(t)]["foot_id"]
        return ms, contact

    def reset_update_swing_trj(self):                               # :685
        self.update_swing_trj = 0
