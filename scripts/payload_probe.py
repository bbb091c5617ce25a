"""BASELINE config 4 (ii): payload variant (k1 = 7) with the TRUE plant mass = 40.05 + m_p while the MPC keeps 40.05: theta_hat_z over time."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cmpc_loader
pkg = cmpc_loader.load()
import torch
planner, com_ref, params, initial = bench.walk_tables()
params = dict(params, N=int(sys.argv[1]) if len(sys.argv) > 1 else 10)
B = 64
mp = np.linspace(0, 10, B)
fleet = pkg.Fleet(B, planner, params, com_ref, initial, k1=7.0, plant_mass=params["mass"] + mp)
rows = []
T = int(sys.argv[2]) if len(sys.argv) > 2 else 190
for t in range(T):
    fleet.step(t)
    if t % 10 == 9 or t == T - 1:
        th = fleet.theta[:, 2].cpu().numpy(); z = fleet.com_pos[:, 2].cpu().numpy(); al = fleet.alive.cpu().numpy()
        rows.append((t, al.mean(), th[[8, 32, 63]].tolist(), z[[8, 32, 63]].tolist()))
        print(t, "alive %.2f" % al.mean(), "theta_z(mp=1.3,5.1,10)", np.round(th[[8, 32, 63]], 3), "expected", np.round(-mp[[8, 32, 63]] * 9.81, 2), "z", np.round(z[[8, 32, 63]], 4), flush=True)
