"""Batched closed loop on the device (`Fleet`, SURVEY.md 8f N1 + N3) against the single-robot drop-in class."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fleet_reproduces_single_robot_loop(pkg):
    import torch
    from oracle.mpc_ref import surrogate_walk
    from oracle.walk import load_walk
    from cmpc_b200.centroidal_mpc_vertices import centroidal_mpc
    t0, t1, B = 0, 290, 8                                   # standing, first single support, landing, step adjustment
    planner, com_ref, params, initial = load_walk()
    mpc = centroidal_mpc(initial, planner, params, com_ref, None, None)
    ref = surrogate_walk(mpc, initial, t0, t1, params["mass"], hw_trace=initial["hw_meas"])
    plan_ref = np.array([s["pos"] for s in planner.plan])
    planner, com_ref, params, initial = load_walk()
    fleet = pkg.Fleet(B, planner, params, com_ref, initial, hw_trace=initial["hw_meas"])
    x0, com, foot, gam = fleet.assemble(250)                # device assembly == host assembly of the drop-in class
    from cmpc_b200.assembly import PlanTables, ReferenceTables, assemble_tick
    cur = {"com": {"pos": initial["com"]["pos"], "vel": initial["com"]["vel"]}, "hw": {"val": initial["hw"]["val"]},
           "lfoot": {"pos": initial["lfoot"]["pos"]}, "rfoot": {"pos": initial["rfoot"]["pos"]}}
    hx0, hcom, hfoot, hgam = assemble_tick(PlanTables(planner.plan), ReferenceTables(com_ref, planner), planner.plan, params, cur, np.zeros(3), 250)
    assert np.array_equal(x0[3].cpu().numpy(), hx0) and np.array_equal(com[3].cpu().numpy(), hcom)
    assert np.array_equal(foot[3].cpu().numpy(), hfoot) and np.array_equal(gam[3].cpu().numpy(), hgam)
    traj = []
    for t in range(t0, t1):
        fleet.step(t)
        traj.append(torch.cat([fleet.com_pos, fleet.com_vel], 1).cpu().numpy())
    traj = np.array(traj)                                   # [ticks, B, 6]
    assert bool(fleet.alive.all())
    assert np.abs(traj - traj[:, :1]).max() == 0.0          # identical robots stay bit-identical
    # (the velocity the plant feeds back carries the push after tick 800 only; none here)
    assert np.abs(traj[:, 0, 0:3] - ref[:, 0:3]).max() <= 1e-7
    assert np.abs(fleet.plan[0].cpu().numpy() - plan_ref).max() <= 1e-7


def test_fleet_disturbance_sweep_runs(pkg):
    import torch
    from oracle.walk import load_walk
    planner, com_ref, params, initial = load_walk()
    B = 256
    fleet = pkg.Fleet(B, planner, params, com_ref, initial, hw_trace=initial["hw_meas"])
    g = torch.Generator(device="cpu").manual_seed(0)
    kick = torch.zeros((B, 3), dtype=torch.float64)
    kick[:, 1] = torch.linspace(0.0, 0.3, B, dtype=torch.float64)          # lateral velocity kick at tick 20, 0 .. 0.3 m/s
    for t in range(0, 60):
        fleet.step(t, push=kick.to(fleet.dev) if t == 20 else None)
    alive = fleet.alive.cpu().numpy()
    assert alive[0] and alive.sum() >= 8                                      # small kicks are absorbed ...
    assert np.all(np.diff(alive.astype(int)) <= 0)                            # ... and survival is monotone in the kick size
