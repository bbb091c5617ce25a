"""Full surrogate walk, GPU drop-in class vs the C oracle driving the same loop: max CoM deviation (north_star: <= 1 mm).
Usage: python scripts/closed_loop_full.py [N] [t_end]   -> one JSON line (copy into profiles/)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import cmpc_loader
pkg = cmpc_loader.load()
from oracle.mpc_ref import surrogate_walk
from oracle.walk import load_walk
from cmpc_b200.centroidal_mpc_vertices import centroidal_mpc
from test_gpu_closed_loop import _oracle_mpc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t_end = int(sys.argv[2]) if len(sys.argv) > 2 else 1970 - N
over = {k: float(v) for k, v in (a.split("=") for a in sys.argv[3:])}       # solver option overrides of the GPU side (experiments)
res = {}
for which in ("gpu", "oracle"):
    planner, com_ref, params, initial = load_walk()
    params["N"] = N
    mpc = centroidal_mpc(initial, planner, params, com_ref, None, None, **over) if which == "gpu" else _oracle_mpc(initial, planner, params, com_ref)
    its = []
    t0 = time.time()
    traj = surrogate_walk(mpc, initial, 0, t_end, params["mass"], hw_trace=initial["hw_meas"],
                          record=lambda t, m, cur: its.append(m.last_iters if which == "gpu" else m.last.iters))
    res[which] = dict(traj=traj, secs=time.time() - t0, iters=float(np.mean(its)), plan=np.array([s["pos"] for s in planner.plan]))
a, b = res["gpu"]["traj"], res["oracle"]["traj"]
print(json.dumps({"N": N, "ticks": int(t_end), "overrides": over, "max_com_pos_dev_m": float(np.abs(a[:, 0:3] - b[:, 0:3]).max()),
                  "max_com_vel_dev": float(np.abs(a[:, 3:6] - b[:, 3:6]).max()), "max_theta_hat_dev": float(np.abs(a[:, 9:12] - b[:, 9:12]).max()),
                  "max_plan_dev_m": float(np.abs(res["gpu"]["plan"] - res["oracle"]["plan"]).max()),
                  "gpu_s_per_tick": res["gpu"]["secs"] / t_end, "oracle_s_per_tick": res["oracle"]["secs"] / t_end,
                  "gpu_iters_per_tick": res["gpu"]["iters"], "oracle_iters_per_tick": res["oracle"]["iters"]}))
