"""ORACLE (test infrastructure) -- the reference `centroidal_mpc` class restated on the oracle solver.

Same constructor / `solve(current, t)` / `model_state` contract as
`code/centroidal_mpc_vertices.py:7,:480-683`, with CasADi/IPOPT replaced by the restated
interior-point oracle (`oracle/ipm_py.py` or the C build `oracle/ipm_c`).  Used by tests to drive
the surrogate closed loop of SURVEY.md section 8d and as the checker for the product class.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import copy

import numpy as np

from . import ipm_py
from .walk import assemble

NS = 52


class centroidal_mpc:  # noqa: N801  (reference class name)
    def __init__(self, initial, footstep_planner, params, CoM_ref, contact_trj_l=None, contact_trj_r=None,
                 solver=None, opts=None, k1=None, k2=None, eps_reg=1e-9):
        self.params = params
        self.N = params["N"]
        self.mass = params["mass"]
        self.g = params["g"]
        self.delta = params["world_time_step"] * params["mpc_rate"]
        self.mpc_rate = params["mpc_rate"]
        self.footstep_planner = footstep_planner
        self.CoM_ref = CoM_ref
        self.k1, self.k2 = k1, k2
        self.eps_reg = eps_reg
        self.update_contact_flag = 0
        self.update_swing_trj = 0
        self.solver = solver or ipm_py.solve
        self.opts = opts or ipm_py.Options.oracle_T()
        self.warm = None                     # primal warm start, unshifted (:630-631)
        self.last = None
        self.last_problem = None
        self.model_state = {"com": {"pos": np.zeros(3), "vel": np.zeros(3), "acc": np.zeros(3)},
                            "hw": {"val": np.zeros(3), "dot": np.zeros(3)},
                            "theta_hat": {"val": np.zeros(3)},
                            "ang_contact_left": {"val": np.zeros(3)}, "pos_contact_left": {"val": np.zeros(3)},
                            "ang_contact_right": {"val": np.zeros(3)}, "pos_contact_right": {"val": np.zeros(3)},
                            "mpc_new_contact": {"val": np.zeros(3)}, "counter": {"val": 0}}

    def solve(self, current, t):
        fp, N = self.footstep_planner, self.N
        prob = assemble(fp, self.CoM_ref, self.params, current, self.model_state["theta_hat"]["val"], t,
                        k1=self.k1, k2=self.k2, eps_reg=self.eps_reg)
        self.last_problem = prob
        res = self.solver(prob, self.warm, self.opts)
        self.last = res
        if res.status != 0:                                    # :605-614 -> the reference crashes here
            raise RuntimeError("oracle solve failed: status %d at t=%d" % (res.status, t))
        X, U = res.X(N), res.U(N)
        self.x, self.u, self.x_collect = X[:, 1].copy(), U[:, 0].copy(), X.copy()    # :614-617
        self.warm = res.w.copy() if getattr(self, "use_warm", True) else None                                                    # :630-631
        gl0, gr0 = prob.gl[0], prob.gr[0]
        Vl = self.u[0:3] + self.u[3:6] + self.u[6:9] + self.u[9:12]
        Vr = self.u[12:15] + self.u[15:18] + self.u[18:21] + self.u[21:24]
        acc = (gl0 * Vl + gr0 * Vr) / self.mass + np.array([0, 0, -self.g])          # :636
        hdot0 = (X[6:9, 1] - X[6:9, 0]) / self.delta                                  # f(x0,u0)[6:9]
        ms = self.model_state
        ms["com"]["pos"], ms["com"]["vel"], ms["com"]["acc"] = self.x[0:3].copy(), self.x[3:6].copy(), acc
        ms["hw"]["val"] = self.x[6:9].copy()
        ms["hw"]["dot"] = 0.01 * hdot0 * self.delta * self.mpc_rate                  # :283,:643
        ms["theta_hat"]["val"] = self.x[9:12].copy()
        ms["ang_contact_left"]["val"], ms["pos_contact_left"]["val"] = self.x[12], self.x[13:16].copy()
        ms["ang_contact_right"]["val"], ms["pos_contact_right"]["val"] = self.x[16], self.x[17:20].copy()
        ms["counter"]["val"] = 0
        if self.params["update_contact"] == "YES":                                   # :656-675
            now = fp.get_phase_at_time(t)
            nxt = fp.get_phase_at_time(t + N * self.mpc_rate - 1)
            if now == "ss" and nxt == "ds" and self.update_contact_flag == 0:
                self.update_contact_flag = 1
                ms["counter"]["val"] = 1
                idx = fp.get_step_index_at_time(t)
                sel = slice(17, 20) if fp.plan[idx]["foot_id"] == "lfoot" else slice(13, 16)
                fp.plan[idx + 1]["pos"] = self.x_collect[sel, N].copy()
                ms["mpc_new_contact"]["val"] = self.x_collect[sel, N].copy()
            if now == "ds":
                self.update_contact_flag = 0
        contact = fp.get_phase_at_time(t)
        if contact == "ss":
            contact = fp.plan[fp.get_step_index_at_time(t)]["foot_id"]
        return ms, contact

    def reset_update_swing_trj(self):
        self.update_swing_trj = 0


def surrogate_walk(mpc, initial, t0, t1, mass, push=True, record=None, verbose=False, hw_trace=None):
    """Closed loop with the centroidal model itself as the plant.

    CoM position / velocity at t+1 := the MPC's own x_1; theta_hat is fed back by the MPC object as in
    :485; the reference's lateral push (3 N on two bodies for 800 < t < 900,
    `code/simulation.py:195-198`) enters as dv_y += 6/m * 0.01 per tick.  The measured whole-body
    angular momentum cannot come from the centroidal model (row :224 would pin it to ~0 for ever and
    the single-support ticks become infeasible); `hw_trace[t]` replays the reference's own recorded
    measurement (`original_code/cuhw.txt`) instead.
    """
    cur = {"com": {"pos": np.array(initial["com"]["pos"], float), "vel": np.array(initial["com"]["vel"], float)},
           "hw": {"val": np.array(initial["hw"]["val"], float)},
           "lfoot": {"pos": np.array(initial["lfoot"]["pos"], float)},
           "rfoot": {"pos": np.array(initial["rfoot"]["pos"], float)}}
    traj = []
    if hw_trace is not None:
        cur["hw"]["val"] = np.array(hw_trace[min(t0, len(hw_trace) - 1)], float)
    for t in range(t0, t1):
        ms, contact = mpc.solve(cur, t)
        if record is not None:
            record(t, mpc, cur)
        traj.append(np.concatenate([ms["com"]["pos"], ms["com"]["vel"], ms["hw"]["val"], ms["theta_hat"]["val"]]))
        cur["com"]["pos"] = ms["com"]["pos"].copy()
        cur["com"]["vel"] = ms["com"]["vel"].copy()
        if push and 800 < t < 900:
            cur["com"]["vel"][1] += 6.0 / mass * 0.01
        cur["hw"]["val"] = ms["hw"]["val"].copy() if hw_trace is None else np.array(hw_trace[min(t + 1, len(hw_trace) - 1)], float)
        cur["lfoot"]["pos"][2] = float(ms["ang_contact_left"]["val"])
        cur["rfoot"]["pos"][2] = float(ms["ang_contact_right"]["val"])
        if ms["counter"]["val"] == 1:
            ms["counter"]["val"] = 0
            mpc.reset_update_swing_trj()
        if verbose:
            print("t %4d %s it %3d cost %.6e" % (t, contact, mpc.last.iters, mpc.last.cost), flush=True)
    return np.array(traj)
