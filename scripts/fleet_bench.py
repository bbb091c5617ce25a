"""Closed-loop fleet throughput: B robots walking the recorded plan with per-robot velocity disturbances, device-resident.
Usage: python scripts/fleet_bench.py [B] [t0] [t1] [N] -> one JSON line (robot-ticks per second, survival)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cmpc_loader
pkg = cmpc_loader.load()
from oracle.walk import load_walk          # fixtures loader only (walk_inputs.npz)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t1 = int(sys.argv[3]) if len(sys.argv) > 3 else 400
N = int(sys.argv[4]) if len(sys.argv) > 4 else 20
planner, com_ref, params, initial = load_walk()
params["N"] = N
fleet = pkg.Fleet(B, planner, params, com_ref, initial, hw_trace=initial["hw_meas"])
g = torch.Generator(device="cuda").manual_seed(0)
its, ms = [], []
torch.cuda.synchronize()
tic = time.perf_counter()
for t in range(t0, t1):
    noise = torch.randn((B, 3), generator=g, device="cuda", dtype=torch.float64) * 2e-4      # 0.2 mm/s per tick, per robot
    out = fleet.step(t, vel_noise=noise)
    if t % 20 == 0:
        its.append(float(out["iters"].double().mean()))
torch.cuda.synchronize()
dt = time.perf_counter() - tic
alive = int(fleet.alive.sum())
print(json.dumps({"robots": B, "N": N, "ticks": t1 - t0, "seconds": dt, "robot_ticks_per_s": B * (t1 - t0) / dt, "ms_per_tick": 1e3 * dt / (t1 - t0),
                  "alive": alive, "mean_iters_sampled": float(np.mean(its)),
                  "com_spread_m": float(fleet.com_pos.std(dim=0).max())}))
