"""Many robots, one GPU: a batched surrogate closed loop with everything resident on the device (SURVEY.md 8f rows N1 + N3).

`Fleet` advances B independent copies of the reference's control loop (`code/simulation.py:193-248` restricted to the
centroidal quantities) one tick at a time:

  1. per-tick parameter assembly of `centroidal_mpc.solve` (MPC file :482-600) as table gathers on the device
     (the reference's O(N * steps) Python loops and 4N `set_value` calls);
  2. one batched solve (`cmpc_solve_device`), warm-started from the previous tick's device-resident iterate;
  3. outputs (:633-649), step adjustment (:656-675) as a batched scatter into the per-robot plan;
  4. the surrogate plant of DESIGN.md section 3 (CoM := predicted x_1, lateral push, replayed angular momentum), plus
     optional per-robot disturbances for robustness sweeps (BASELINE config 3).

torch is plumbing here (device tensors, gathers); the solve is the CUDA kernel behind the C ABI.
"""
from __future__ import annotations

import numpy as np

from ._lib import COLD, WARM_FULL, BatchSolver
from .assembly import PlanTables, ReferenceTables


class Fleet:
    def __init__(self, batch, footstep_planner, params, CoM_ref, initial, hw_trace=None, device=0, k1=None, mass=None,
                 **solver_overrides):
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
        T = min(min(len(c) for c in rf.com), len(rf.pos_l))
        self.T = T
        self.com_tab = torch.as_tensor(np.stack([c[:T] for c in rf.com], axis=1), **f64)                    # [T, 9]
        self.foot_tab = torch.as_tensor(np.concatenate([rf.pos_l[:T], rf.pos_r[:T], rf.yaw_l[:T, None], rf.yaw_r[:T, None]], 1), **f64)
        self.gamma_tab = torch.as_tensor(tb.gamma, **f64)                                                  # [T_plan, 2]
        self.plan = torch.as_tensor(np.stack([s["pos"] for s in footstep_planner.plan]), **f64).repeat(self.B, 1, 1).contiguous()
        self.first_swing_left = params["first_swing"] == "lfoot"
        delta = params["world_time_step"] * self.rate
        self.mass = torch.full((self.B,), float(params["mass"]), **f64) if mass is None else torch.as_tensor(mass, **f64).contiguous()
        k1v = (5.0 if self.rate == 10 else 4.0) if k1 is None else k1
        self.k1 = torch.as_tensor(np.broadcast_to(np.asarray(k1v, float), (self.B,)).copy(), **f64)
        self.solver = BatchSolver(self.N, self.B, device=device, delta=delta, grav=params["g"], w_rate=0.0 if self.rate == 10 else 1.0,
                                  **solver_overrides)
        rep = lambda v: torch.as_tensor(np.asarray(v, float), **f64).repeat(self.B, 1).contiguous()
        self.com_pos, self.com_vel = rep(initial["com"]["pos"]), rep(initial["com"]["vel"])
        self.hw = rep(initial["hw"]["val"])
        self.theta = torch.zeros((self.B, 3), **f64)
        self.yaw = torch.as_tensor([[initial["lfoot"]["pos"][2], initial["rfoot"]["pos"][2]]], **f64).repeat(self.B, 1).contiguous()
        self.hw_trace = None if hw_trace is None else torch.as_tensor(np.asarray(hw_trace, float), **f64)
        self.update_flag = False
        self.alive = torch.ones(self.B, dtype=torch.bool, device=self.dev)
        self.first = True
        self.g = float(params["g"])
        self.out = None

    # ---- (1) parameter assembly on the device
    def assemble(self, t: int):
        torch, N, rate, B = self.torch, self.N, self.rate, self.B
        tt = t + (1 + torch.arange(N, device=self.dev)) * rate
        if int(tt[-1]) >= self.T:
            raise IndexError("tick %d: the horizon runs past the reference tables (%d rows)" % (t, self.T))      # :567
        x0 = torch.empty((B, 20), dtype=torch.float64, device=self.dev)
        x0[:, 0:3], x0[:, 3:6], x0[:, 6:9], x0[:, 9:12] = self.com_pos, self.com_vel, self.hw, self.theta
        x0[:, 12], x0[:, 16] = self.yaw[:, 0], self.yaw[:, 1]
        if t < 200:                                                                                            # :493-495
            x0[:, 13:16], x0[:, 17:20] = self.foot_tab[t, 0:3], self.foot_tab[t, 3:6]
        else:                                                                                                  # :496-503
            index = self.tables.step_index_at(t - 70)
            a, b = index + (index % 2), index + (index - 1) % 2
            il, ir = (a, b) if self.first_swing_left else (b, a)
            x0[:, 13:16], x0[:, 17:20] = self.plan[:, il], self.plan[:, ir]
        com = self.com_tab[tt].unsqueeze(0).expand(B, N, 9).contiguous()
        foot = self.foot_tab[tt].clone()
        tq = t + (1 + torch.arange(N, device=self.dev) // 3) * rate                                            # yaw quirk :599-600
        foot[:, 6:8] = self.foot_tab[tq, 6:8]
        foot = foot.unsqueeze(0).expand(B, N, 8).contiguous()
        tg = t + rate * torch.arange(N + 1, device=self.dev)
        if int(tg[-1]) >= self.gamma_tab.shape[0]:
            raise TypeError("horizon reaches beyond the footstep plan")
        gamma = self.gamma_tab[tg].unsqueeze(0).expand(B, N + 1, 2).contiguous()
        return x0, com, foot, gamma

    # ---- (2)-(4) one control tick for the whole fleet
    def step(self, t: int, push=None, vel_noise=None):
        """push: optional [B, 3] velocity increment applied after the tick (robustness sweeps);
        vel_noise: optional [B, 3] measurement disturbance added to the CoM velocity fed to the next tick."""
        torch = self.torch
        x0, com, foot, gamma = self.assemble(t)
        mode = COLD if self.first else WARM_FULL
        self.out = self.solver.solve_device(x0, com, foot, gamma, self.mass, self.k1, mode, out=self.out)
        self.first = False
        out = self.out
        ok = out["status"] == 0
        self.alive &= ok                                       # a failed solve is where the reference would crash (:605-614)
        x1, u0, xN = out["x1"], out["u0"], out["xN"]
        upd = self.alive.unsqueeze(1)
        self.com_pos = torch.where(upd, x1[:, 0:3], self.com_pos)
        self.com_vel = torch.where(upd, x1[:, 3:6], self.com_vel)
        self.theta = torch.where(upd, x1[:, 9:12], self.theta)
        self.yaw = torch.where(upd, torch.stack([x1[:, 12], x1[:, 16]], 1), self.yaw)
        if self.hw_trace is not None:
            self.hw = self.hw_trace[min(t + 1, self.hw_trace.shape[0] - 1)].unsqueeze(0).expand(self.B, 3).contiguous()
        else:
            self.hw = torch.where(upd, x1[:, 6:9], self.hw)
        if 800 < t < 900:                                      # the reference's push, simulation.py:195-198
            self.com_vel[:, 1] += 6.0 / self.mass * 0.01
        if push is not None:
            self.com_vel += push
        if vel_noise is not None:
            self.com_vel += vel_noise
        tb = self.tables
        if self.params["update_contact"] == "YES":             # :656-675, same tick for every robot -> a batched scatter
            now, after = tb.phase_at(t), tb.phase_at(t + self.N * self.rate - 1)
            if now == "ss" and after == "ds" and not self.update_flag:
                self.update_flag = True
                idx = tb.step_index_at(t)
                sl = slice(17, 20) if tb.left_support[idx] else slice(13, 16)
                self.plan[:, idx + 1] = torch.where(upd, xN[:, sl], self.plan[:, idx + 1])
            if now == "ds":
                self.update_flag = False
        return out

    def com_acc(self, gamma0):
        u0 = self.out["u0"]
        Vl = u0[:, 0:12].reshape(-1, 4, 3).sum(1)
        Vr = u0[:, 12:24].reshape(-1, 4, 3).sum(1)
        acc = (gamma0[0] * Vl + gamma0[1] * Vr) / self.mass.unsqueeze(1)
        acc[:, 2] -= self.g
        return acc                                             # :636
