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


def test_device_assembly_against_the_literal_restatement(pkg):
    """cmpc_assemble_device (one gather kernel, every robot at its own tick, per-robot step-adjusted plans) against
    oracle/walk.assemble, the loop-by-loop restatement of MPC file :482-600 -- bit-exact, including the yaw quirk, the
    first_swing parity rule and the error codes for horizons that leave the tables."""
    import torch
    from oracle.walk import assemble, load_walk
    planner, com_ref, params, initial = load_walk()
    planner.position_contacts_ref["contact_left"][:, 2] = 0.01 * np.arange(len(planner.position_contacts_ref["contact_left"]))   # non-trivial yaws
    planner.position_contacts_ref["contact_right"][:, 2] = -0.02 * np.arange(len(planner.position_contacts_ref["contact_right"]))
    rng = np.random.default_rng(5)
    for N in (10, 20):
        params["N"] = N
        ticks = np.array([0, 1, 57, 199, 200, 201, 269, 270, 271, 299, 300, 805, 1500, 1970 - N, 1971 - N, 1990, 5, 640], np.int64)
        B = len(ticks)
        fleet = pkg.Fleet(B, planner, params, com_ref, initial, tick_offset=ticks)
        st = {k: rng.normal(size=(B, 3)) for k in ("pos", "vel", "hw", "th")}
        yaw = rng.normal(size=(B, 2))
        plan = np.stack([s["pos"] for s in planner.plan])[None] + 1e-3 * rng.normal(size=(B, len(planner.plan), 3))
        t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=fleet.dev)
        fleet.com_pos, fleet.com_vel, fleet.hw, fleet.theta, fleet.yaw, fleet.plan = t64(st["pos"]), t64(st["vel"]), t64(st["hw"]), t64(st["th"]), t64(yaw), t64(plan)
        x0, com, foot, gam = [a.cpu().numpy() for a in fleet.assemble(0)]
        err = fleet._asm[4].cpu().numpy()
        for b, t in enumerate(ticks):
            cur = {"com": {"pos": st["pos"][b], "vel": st["vel"][b]}, "hw": {"val": st["hw"][b]},
                   "lfoot": {"pos": np.array([0, 0, yaw[b, 0], 0, 0, 0.0])}, "rfoot": {"pos": np.array([0, 0, yaw[b, 1], 0, 0, 0.0])}}
            for j, s_ in enumerate(planner.plan):
                s_["pos"] = plan[b, j].copy()
            try:
                ref = assemble(planner, com_ref, params, cur, st["th"][b], int(t))
            except (IndexError, TypeError) as e:
                assert err[b] == (1 if isinstance(e, IndexError) else 2), (N, int(t), err[b], e)
                continue
            assert err[b] == 0, (N, int(t))
            assert np.array_equal(x0[b], ref.x0), (N, int(t))
            assert np.array_equal(com[b], ref.com_ref.T) and np.array_equal(gam[b], np.stack([ref.gl, ref.gr], 1))
            assert np.array_equal(foot[b], np.concatenate([ref.pl_ref, ref.pr_ref, ref.al_ref[None], ref.ar_ref[None]], 0).T), (N, int(t))


def test_fleet_against_the_oracle_driven_loop(pkg):
    """N3: robots at DIFFERENT phases of the walk in one batch (tick offsets 0 and 180) stay within 1 mm of the loops
    driven by the C oracle from the same initial states (standing start; first lift-off, landing and step adjustment)."""
    import torch
    from oracle.mpc_ref import surrogate_walk
    from oracle.walk import load_walk
    from test_gpu_closed_loop import _oracle_mpc
    N, T = 10, 120
    offs = [0, 180]
    refs, plans = [], []
    for o in offs:
        planner, com_ref, params, initial = load_walk()
        params["N"] = N
        initial["com"]["pos"] = np.array([com_ref["pos_x"][o], com_ref["pos_y"][o], 0.72])
        initial["com"]["vel"] = np.array([com_ref["vel_x"][o], com_ref["vel_y"][o], 0.0])
        refs.append(surrogate_walk(_oracle_mpc(initial, planner, params, com_ref), initial, o, o + T, params["mass"], hw_trace=initial["hw_meas"]))
        plans.append(np.array([s["pos"] for s in planner.plan]))
    planner, com_ref, params, initial = load_walk()
    params["N"] = N
    fleet = pkg.Fleet(len(offs), planner, params, com_ref, initial, hw_trace=initial["hw_meas"], tick_offset=offs)
    f64 = dict(dtype=torch.float64, device=fleet.dev)
    fleet.com_pos = torch.as_tensor(np.array([[com_ref["pos_x"][o], com_ref["pos_y"][o], 0.72] for o in offs]), **f64)
    fleet.com_vel = torch.as_tensor(np.array([[com_ref["vel_x"][o], com_ref["vel_y"][o], 0.0] for o in offs]), **f64)
    fleet.hw = torch.as_tensor(np.array([initial["hw_meas"][o] for o in offs]), **f64)
    traj = []
    for t in range(T):
        fleet.step(t)
        traj.append(fleet.com_pos.cpu().numpy())
    traj = np.array(traj)
    assert bool(fleet.alive.all())
    for b in range(len(offs)):
        assert np.abs(traj[:, b] - refs[b][:, 0:3]).max() <= 1e-3, (b, np.abs(traj[:, b] - refs[b][:, 0:3]).max())
        assert np.abs(fleet.plan[b].cpu().numpy() - plans[b]).max() <= 1e-3


def test_payload_sweep_with_true_plant_mass(pkg):
    """BASELINE config 4 (ii): payload variant (k1 = 7) on robots whose TRUE mass is 40.05 + m_p while the MPC keeps 40.05
    (the reference never changes `mass`, it lets theta_hat adapt).  The plant is integrated from the applied
    contact forces.  Checked: every robot keeps solving; the CoM sags monotonically with the payload; theta_hat_z grows
    negative monotonically with the payload and equals the estimator law theta+ = theta + (d/m)(k1 (p - p_ref) + v - v_ref)
    (MPC file :459) integrated along the recorded CoM trajectory.  (The law's gain is d/m per tick: theta_hat_z would need
    ~1e6 ticks to reach -m_p g; the test pins the dynamics, not the limit.)"""
    import torch
    from oracle.walk import load_walk
    planner, com_ref, params, initial = load_walk()
    B, T, k1 = 32, 120, 7.0
    mp = np.linspace(0.0, 10.0, B)
    fleet = pkg.Fleet(B, planner, params, com_ref, initial, k1=k1, plant_mass=params["mass"] + mp)
    th = np.zeros(B)
    for t in range(T):
        pz, vz = fleet.com_pos[:, 2].cpu().numpy(), fleet.com_vel[:, 2].cpu().numpy()
        fleet.step(t)
        th += 0.01 / params["mass"] * (k1 * (pz - com_ref["pos_z"][t + 1]) + vz - com_ref["vel_z"][t + 1])      # reference column 0 of tick t = table row t + 1
    assert bool(fleet.alive.all())
    z = fleet.com_pos[:, 2].cpu().numpy(); thz = fleet.theta[:, 2].cpu().numpy()
    assert np.all(np.diff(z) < 0) and 0.02 < z[0] - z[-1] < 0.08            # heavier payload, lower CoM (a few centimetres at 10 kg)
    assert np.all(np.diff(thz) < 0) and thz[-1] < -1e-3
    assert np.abs(thz - th).max() <= 1e-9 + 1e-6 * np.abs(th).max()
