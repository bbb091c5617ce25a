"""ORACLE (test infrastructure) -- literal restatement of the per-tick parameter assembly.

Follows `centroidal_mpc.solve` in `code/centroidal_mpc_vertices.py:482-600` loop by loop (x0 with
the foot positions overwritten from the plan, contact schedule, CoM / foot references including
the column-major yaw quirk at :563,:583,:599) and the planner queries of
`code/footstep_planner_vertices.py:82-103`.  Inputs come from `tests/golden/walk_inputs.npz`,
which was produced by importing the reference's own planner (tests/golden/make_walk_inputs.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import os

import numpy as np

from .ipm_py import Problem

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class PlanTable:
    """Stand-in for the reference FootstepPlanner restricted to what `solve` touches."""

    def __init__(self, data):
        self.plan = []
        for j in range(len(data["plan_ss"])):
            self.plan.append({"pos": np.array(data["plan_pos"][j], float),
                              "ang": np.array(data["plan_ang"][j], float),
                              "ss_duration": int(data["plan_ss"][j]),
                              "ds_duration": int(data["plan_ds"][j]),
                              "foot_id": "lfoot" if int(data["plan_foot"][j]) == 0 else "rfoot"})
        self.position_contacts_ref = {"contact_left": np.array(data["contact_left"], float),
                                      "contact_right": np.array(data["contact_right"], float)}

    def get_step_index_at_time(self, time):          # footstep_planner_vertices.py:82-88
        t = 0
        for i in range(len(self.plan)):
            t += self.plan[i]["ss_duration"] + self.plan[i]["ds_duration"]
            if t > time:
                return i
        return None

    def get_start_time(self, step_index):            # :90-94
        t = 0
        for i in range(step_index):
            t += self.plan[i]["ss_duration"] + self.plan[i]["ds_duration"]
        return t

    def get_phase_at_time(self, time):               # :96-103
        step_index = self.get_step_index_at_time(time)
        time_in_step = time - self.get_start_time(step_index)
        return "ss" if time_in_step < self.plan[step_index]["ss_duration"] else "ds"


def load_walk(path: str = None):
    data = np.load(path or os.path.join(GOLDEN, "walk_inputs.npz"))
    planner = PlanTable(data)
    com_ref = {k[4:]: np.array(data[k], float) for k in data.files if k.startswith("ref_")}
    params = {"g": 9.81, "h": 0.72, "foot_size": 0.1, "step_height": 0.02, "world_time_step": 0.01,
              "ss_duration": 70, "ds_duration": 30, "first_swing": "rfoot", "µ": 0.5, "N": 10,
              "mass": float(data["mass"]), "update_contact": "YES", "mpc_rate": 1}
    params["eta"] = float(np.sqrt(params["g"] / params["h"]))
    initial = {"lfoot": {"pos": np.array(data["lfoot0"], float)}, "rfoot": {"pos": np.array(data["rfoot0"], float)},
               "com": {"pos": np.array([0.0, 0.0, 0.72]), "vel": np.zeros(3)}, "hw": {"val": np.zeros(3)}}
    initial["hw_meas"] = np.array(data["hw_meas"], float) if "hw_meas" in data.files else None
    return planner, com_ref, params, initial


def assemble(planner, com_ref, params, current, theta_hat, t, k1=None, k2=None, eps_reg=1e-9) -> Problem:
    """x0, gamma, references of one tick -- MPC file :482-600, restated with the same loops."""
    N, rate = params["N"], params["mpc_rate"]
    x0 = np.array([current["com"]["pos"][0], current["com"]["pos"][1], current["com"]["pos"][2],
                   current["com"]["vel"][0], current["com"]["vel"][1], current["com"]["vel"][2],
                   current["hw"]["val"][0], current["hw"]["val"][1], current["hw"]["val"][2],
                   theta_hat[0], theta_hat[1], theta_hat[2],
                   current["lfoot"]["pos"][2], current["lfoot"]["pos"][3], current["lfoot"]["pos"][4], 0.0,
                   current["rfoot"]["pos"][2], current["rfoot"]["pos"][3], current["rfoot"]["pos"][4], 0.0])
    pos_l = planner.position_contacts_ref["contact_left"][:, 3:6]
    pos_r = planner.position_contacts_ref["contact_right"][:, 3:6]
    yaw_l = planner.position_contacts_ref["contact_left"][:, 2]
    yaw_r = planner.position_contacts_ref["contact_right"][:, 2]
    if t < 200:                                                            # :493-503
        cl, cr = pos_l[t], pos_r[t]
    else:
        index = planner.get_step_index_at_time(t - 70)
        if params["first_swing"] == "lfoot":
            cl = planner.plan[index + (index % 2)]["pos"]
            cr = planner.plan[index + (index - 1) % 2]["pos"]
        else:
            cl = planner.plan[index + (index - 1) % 2]["pos"]
            cr = planner.plan[index + (index % 2)]["pos"]
    x0[13:16] = cl
    x0[17:20] = cr
    gl, gr = np.zeros(N + 1), np.zeros(N + 1)
    for i in range(N + 1):                                                 # :517-531
        if planner.get_phase_at_time(t + i * rate) == "ds":
            gl[i], gr[i] = 1.0, 1.0
        else:
            foot = planner.plan[planner.get_step_index_at_time(t + i * rate)]["foot_id"]
            gl[i], gr[i] = (1.0, 0.0) if foot == "lfoot" else (0.0, 1.0)
    ref = np.zeros((9, N))
    pl, pr = np.zeros((3, N)), np.zeros((3, N))
    al3, ar3 = np.zeros((3, N)), np.zeros((3, N))
    keys = ["pos_x", "pos_y", "pos_z", "vel_x", "vel_y", "vel_z", "acc_x", "acc_y", "acc_z"]
    for i in range(N):                                                     # :565-584
        tt = t + (1 + i) * rate
        for j, k in enumerate(keys):
            ref[j, i] = com_ref[k][tt]
        pl[:, i] = pos_l[tt]
        pr[:, i] = pos_r[tt]
        al3[:, i] = yaw_l[tt]
        ar3[:, i] = yaw_r[tt]
    # :599-600 index the 3xN DM with ONE index -> column-major linear index i -> yaw of step i//3
    al = np.array([al3.flatten(order="F")[i] for i in range(N)])
    ar = np.array([ar3.flatten(order="F")[i] for i in range(N)])
    if k1 is None:                                                         # :27-31
        k1, k2 = (5.0, 0.2) if rate == 10 else (4.0, 0.1)
    return Problem(N=N, x0=x0, com_ref=ref, pl_ref=pl, pr_ref=pr, al_ref=al, ar_ref=ar, gl=gl, gr=gr,
                   mass=params["mass"], k1=k1, k2=k2, delta=params["world_time_step"] * rate,
                   grav=params["g"], w_rate=0.0 if rate == 10 else 1.0, eps_reg=eps_reg)
