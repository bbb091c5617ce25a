"""Many robots, one GPU: a batched surrogate closed loop with everything resident on the device (SURVEY.md 8f rows N1 + N3).

`Fleet` advances B independent copies of the reference's control loop (`code/simulation.py:193-248` restricted to the
centroidal quantities) one tick at a time:

  1. per-tick parameter assembly of `centroidal_mpc.solve` (MPC file :482-600) by ONE gather kernel behind the C ABI
     (`cmpc_assemble_device`: the reference's O(N * steps) Python loops and 4N `set_value` calls), each robot at its own tick;
  2. one batched solve (`cmpc_solve_device`), warm-started from the previous tick's device-resident iterate;
  3. outputs (:633-649), step adjustment (:656-675) as a batched scatter into the per-robot plan;
  4. the surrogate plant of DESIGN.md section 3 (CoM := predicted x_1, lateral push, replayed angular momentum), plus
     optional per-robot disturbances for robustness sweeps (BASELINE config 3).

torch is plumbing here (device tensors, gathers); the solve is the CUDA kernel behind the C ABI.
"""
from __future__ import annotations

import numpy as np

from ._lib import COLD, WARM_AUTO, BatchSolver, WalkTables
from .assembly import PlanTables, ReferenceTables


class Fleet:
    def __init__(self, batch, footstep_planner, params, CoM_ref, initial, hw_trace=None, device=0, k1=None, mass=None,
                 tick_offset=None, warm_mode=WARM_AUTO, plant_mass=None, **solver_overrides):
        """tick_offset: optional int array [B]: robot b is at tick t + tick_offset[b] when the fleet is stepped at tick t
        (robots at different phases of the same walk in one batch).
        plant_mass: optional [B] true masses of the simulated robots (BASELINE config 4 (ii): the MPC keeps `mass`, the plant
        carries a payload): the CoM is then integrated from the applied contact forces, p+ = p + d v, v+ = v + d (g + F / m_plant),
        instead of being set to the MPC's own prediction, and theta_hat has something to estimate."""
        import torch
        self.torch = torch
        self.B, self.N = int(batch), int(params["N"])
        self.rate = int(params["mpc_rate"])
        self.params = params
        self.dev = torch.device("cuda", device)
        f64 = dict(dtype=torch.float64, device=self.dev)
        tb = PlanTables(footstep_planner.plan)
        rf = ReferenceTables(CoM_ref, footstep_planner)
        self.tables = tb
        self.first_swing_left = params["first_swing"] == "lfoot"
        self.walk = WalkTables(tb, rf, self.N, self.rate, self.first_swing_left, len(footstep_planner.plan), device=device)
        self.T = self.walk.T_ref
        self.plan = torch.as_tensor(np.stack([s["pos"] for s in footstep_planner.plan]), **f64).repeat(self.B, 1, 1).contiguous()
        delta = params["world_time_step"] * self.rate
        self.mass = torch.full((self.B,), float(params["mass"]), **f64) if mass is None else torch.as_tensor(mass, **f64).contiguous()
        k1v = (5.0 if self.rate == 10 else 4.0) if k1 is None else k1
        self.k1 = torch.as_tensor(np.broadcast_to(np.asarray(k1v, float), (self.B,)).copy(), **f64)
        self.solver = BatchSolver(self.N, self.B, device=device, delta=delta, grav=params["g"], w_rate=0.0 if self.rate == 10 else 1.0,
                                  **solver_overrides)
        self.warm_mode = warm_mode
        self.plant_mass = None if plant_mass is None else torch.as_tensor(np.asarray(plant_mass, float), **f64).contiguous()
        self.delta = delta
        rep = lambda v: torch.as_tensor(np.asarray(v, float), **f64).repeat(self.B, 1).contiguous()
        self.com_pos, self.com_vel = rep(initial["com"]["pos"]), rep(initial["com"]["vel"])
        self.hw = rep(initial["hw"]["val"])
        self.theta = torch.zeros((self.B, 3), **f64)
        self.yaw = torch.as_tensor([[initial["lfoot"]["pos"][2], initial["rfoot"]["pos"][2]]], **f64).repeat(self.B, 1).contiguous()
        self.hw_trace = None if hw_trace is None else torch.as_tensor(np.asarray(hw_trace, float), **f64)
        off = np.zeros(self.B, np.int64) if tick_offset is None else np.asarray(tick_offset, np.int64)
        self.offset = torch.as_tensor(off, device=self.dev)
        self.update_flag = torch.zeros(self.B, dtype=torch.bool, device=self.dev)
        self.alive = torch.ones(self.B, dtype=torch.bool, device=self.dev)
        self.first = True
        self.g = float(params["g"])
        self.out = None
        self._asm = None

    # ---- (1) parameter assembly on the device: one gather kernel behind the C ABI (cmpc_assemble_device)
    def assemble(self, t: int):
        torch = self.torch
        tick = (self.offset + int(t)).to(torch.int32)
        self._asm = self.walk.assemble(tick, self.com_pos, self.com_vel, self.hw, self.theta, self.yaw, self.plan, out=self._asm)
        return self._asm[:4]

    # ---- (2)-(4) one control tick for the whole fleet
    def step(self, t: int, push=None, vel_noise=None):
        """push: optional [B, 3] velocity increment applied after the tick (robustness sweeps);
        vel_noise: optional [B, 3] measurement disturbance added to the CoM velocity fed to the next tick."""
        torch = self.torch
        x0, com, foot, gamma = self.assemble(t)
        mode = COLD if self.first else self.warm_mode
        self.out = self.solver.solve_device(x0, com, foot, gamma, self.mass, self.k1, mode, out=self.out)
        self.first = False
        out = self.out
        tb = self.walk
        tick = self.offset + int(t)
        in_range = self._asm[4] == 0                           # a horizon past the tables is where the reference raises IndexError (:567)
        ok = (out["status"] == 0) & in_range
        self.alive &= ok                                       # a failed solve is where the reference would crash (:605-614)
        x1, u0, xN = out["x1"], out["u0"], out["xN"]
        upd = self.alive.unsqueeze(1)
        if self.plant_mass is None:                            # surrogate plant of DESIGN.md section 3: the model itself
            new_pos, new_vel = x1[:, 0:3], x1[:, 3:6]
        else:                                                  # a plant of another mass under the applied forces (Euler step, :187-190)
            g0 = gamma[:, 0]
            F = g0[:, 0:1] * u0[:, 0:12].reshape(-1, 4, 3).sum(1) + g0[:, 1:2] * u0[:, 12:24].reshape(-1, 4, 3).sum(1)
            acc = F / self.plant_mass.unsqueeze(1)
            acc[:, 2] -= self.g
            new_pos, new_vel = self.com_pos + self.delta * self.com_vel, self.com_vel + self.delta * acc
        self.com_pos = torch.where(upd, new_pos, self.com_pos)
        self.com_vel = torch.where(upd, new_vel, self.com_vel)
        self.theta = torch.where(upd, x1[:, 9:12], self.theta)
        self.yaw = torch.where(upd, torch.stack([x1[:, 12], x1[:, 16]], 1), self.yaw)
        if self.hw_trace is not None:
            self.hw = self.hw_trace[torch.clamp(tick + 1, max=self.hw_trace.shape[0] - 1)].contiguous()
        else:
            self.hw = torch.where(upd, x1[:, 6:9], self.hw)
        pushed = ((tick > 800) & (tick < 900)).to(torch.float64)           # the reference's push, simulation.py:195-198
        self.com_vel[:, 1] += pushed * 6.0 / self.mass * 0.01
        if push is not None:
            self.com_vel += push
        if vel_noise is not None:
            self.com_vel += vel_noise
        if self.params["update_contact"] == "YES":             # :656-675 per robot -> a batched scatter into the per-robot plans
            tc = torch.clamp(tick, max=tb.T_plan - 1)
            ta = torch.clamp(tick + self.N * self.rate - 1, max=tb.T_plan - 1)
            now_ss, after_ds = tb.is_ss[tc], ~tb.is_ss[ta]
            cond = now_ss & after_ds & ~self.update_flag & self.alive
            idx = tb.step_index[tc].long()
            left = tb.left_support[idx]
            sel = torch.where(left.unsqueeze(1), xN[:, 17:20], xN[:, 13:16])
            rows = torch.nonzero(cond).squeeze(1)
            if rows.numel():
                self.plan[rows, idx[rows] + 1] = sel[rows]
            self.update_flag = (self.update_flag | cond) & now_ss           # (:673-675: the flag is cleared in double support)
        return out

    def com_acc(self, gamma0):
        u0 = self.out["u0"]
        Vl = u0[:, 0:12].reshape(-1, 4, 3).sum(1)
        Vr = u0[:, 12:24].reshape(-1, 4, 3).sum(1)
        acc = (gamma0[0] * Vl + gamma0[1] * Vr) / self.mass.unsqueeze(1)
        acc[:, 2] -= self.g
        return acc                                             # :636
