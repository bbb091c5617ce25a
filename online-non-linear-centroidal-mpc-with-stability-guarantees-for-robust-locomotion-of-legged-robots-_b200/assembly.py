"""Per-tick parameter assembly of `centroidal_mpc.solve` (reference `code/centroidal_mpc_vertices.py:482-600`),
vectorised: the reference's O(N * steps) Python loops and 4N scalar `set_value` calls become table look-ups.

Restated behaviour (file:line of the reference):
  * x0 = [CoM pos, CoM vel, h_w, theta_hat of the previous solution, yaw_l, p_l, yaw_r, p_r] with the foot
    positions overwritten from the plan (:482-509): t < 200 -> contact reference table row t, else
    plan[index + (index-1)%2] / plan[index + index%2] with index = step_index(t - 70) (first_swing parity);
  * gamma_l/r[i] for i = 0..N from phase(t + i*rate) and plan[step_index]['foot_id'] (:515-534);
  * CoM / foot references sampled at t + (1+i)*rate (:565-584);
  * yaw references: the reference fills a 3xN DM whose rows all hold the yaw and then reads it with ONE index,
    which is column-major -> stage i receives the yaw of horizon step i // 3 (:563, :583, :599-600).  Kept.
Planner queries follow `code/footstep_planner_vertices.py:82-103`.
"""
from __future__ import annotations

import numpy as np

_REF_KEYS = ("pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z", "acc_x", "acc_y", "acc_z")


class PlanTables:
    """Phase / step-index / contact tables over absolute ticks, built once from the plan's durations."""

    def __init__(self, plan):
        ss = np.array([int(s["ss_duration"]) for s in plan])
        ds = np.array([int(s["ds_duration"]) for s in plan])
        self.ss = ss
        self.ends = np.cumsum(ss + ds)
        self.starts = self.ends - (ss + ds)
        self.left_support = np.array([s["foot_id"] == "lfoot" for s in plan])
        T = int(self.ends[-1])
        t = np.arange(T)
        self.step_index = np.searchsorted(self.ends, t, side="right")
        self.is_ss = (t - self.starts[self.step_index]) < ss[self.step_index]
        sup_left = self.left_support[self.step_index]
        self.gamma = np.ones((T, 2))
        self.gamma[self.is_ss & ~sup_left, 0] = 0.0     # right foot supports -> left swings
        self.gamma[self.is_ss & sup_left, 1] = 0.0
        self.T = T

    def step_index_at(self, time: int) -> int:
        if time < 0:
            return 0                                      # cumulative sum is > any negative time at i = 0
        if time >= self.T:
            raise TypeError("time %d is beyond the footstep plan (the reference returns None here)" % time)
        return int(self.step_index[time])

    def phase_at(self, time: int) -> str:
        i = self.step_index_at(time)
        return "ss" if (time - self.starts[i]) < self.ss[i] else "ds"

    def gamma_at(self, times):
        times = np.asarray(times)
        if times.max() >= self.T:
            raise TypeError("horizon reaches beyond the footstep plan")
        return self.gamma[times]


class ReferenceTables:
    """CoM reference (9 x T, ragged lengths kept) and contact reference tables as arrays."""

    def __init__(self, CoM_ref, footstep_planner):
        self.com = [np.asarray(CoM_ref[k], float).ravel() for k in _REF_KEYS]
        cl = np.asarray(footstep_planner.position_contacts_ref["contact_left"], float)
        cr = np.asarray(footstep_planner.position_contacts_ref["contact_right"], float)
        self.pos_l, self.pos_r = cl[:, 3:6], cr[:, 3:6]          # :80-81
        self.yaw_l, self.yaw_r = cl[:, 2], cr[:, 2]              # :83-84


def assemble_tick(tables: PlanTables, refs: ReferenceTables, plan, params, current, theta_hat, t: int):
    """Returns (x0[20], com_ref[N,9], foot_ref[N,8], gamma[N+1,2]) of tick t."""
    N, rate = int(params["N"]), int(params["mpc_rate"])
    x0 = np.empty(20)
    x0[0:3] = current["com"]["pos"][0:3]
    x0[3:6] = current["com"]["vel"][0:3]
    x0[6:9] = current["hw"]["val"][0:3]
    x0[9:12] = theta_hat[0:3]
    x0[12] = current["lfoot"]["pos"][2]
    x0[16] = current["rfoot"]["pos"][2]
    if t < 200:                                                    # :493-495
        x0[13:16], x0[17:20] = refs.pos_l[t], refs.pos_r[t]
    else:                                                          # :496-503
        index = tables.step_index_at(t - 70)
        a, b = index + (index % 2), index + (index - 1) % 2
        if params["first_swing"] == "lfoot":
            x0[13:16], x0[17:20] = plan[a]["pos"], plan[b]["pos"]
        else:
            x0[13:16], x0[17:20] = plan[b]["pos"], plan[a]["pos"]
    gamma = tables.gamma_at(t + rate * np.arange(N + 1)).copy()    # :517-531
    tt = t + (1 + np.arange(N)) * rate                             # :567
    com = np.empty((N, 9))
    for j in range(9):
        com[:, j] = refs.com[j][tt]                                # IndexError past the table, as the reference
    foot = np.empty((N, 8))
    foot[:, 0:3] = refs.pos_l[tt]
    foot[:, 3:6] = refs.pos_r[tt]
    tq = t + (1 + np.arange(N) // 3) * rate                        # the column-major yaw quirk (:599-600)
    foot[:, 6] = refs.yaw_l[tq]
    foot[:, 7] = refs.yaw_r[tq]
    return x0, com, foot, gamma


def pack_instances(instances):
    """Stack a list of (x0, com_ref, foot_ref, gamma) tuples into the instance-major batch arrays of the C ABI."""
    x0 = np.stack([i[0] for i in instances])
    com = np.stack([i[1] for i in instances])
    foot = np.stack([i[2] for i in instances])
    gam = np.stack([i[3] for i in instances])
    return x0, com, foot, gam
