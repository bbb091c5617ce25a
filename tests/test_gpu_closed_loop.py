"""Closed loop on the GPU: the product's drop-in `centroidal_mpc` class drives the surrogate walk of SURVEY.md 8d and
must keep the CoM within 1 mm of the walk driven by the oracle (BASELINE.json north_star).  The plant is the
centroidal model itself; the measured angular momentum is the reference's own recording (original_code/cuhw.txt,
stored in tests/golden/walk_inputs.npz)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_mpc(initial, planner, params, com_ref, k1=None):
    from oracle import ipm_c
    from oracle.mpc_ref import centroidal_mpc as RefMPC

    class R:
        pass

    def solver(prob, w0, opts):
        import hostsim  # only for pack(): layout helper of the tests
        x0, com, foot, gam = hostsim.pack(prob)
        warm = None
        if w0 is not None:
            N = prob.N
            warm = (np.stack([w0[52 * i:52 * i + 20] for i in range(N + 1)]), np.stack([w0[52 * i + 20:52 * i + 52] for i in range(N)]))
        r = ipm_c.solve_packed(prob.N, x0, com, foot, gam, prob.mass, prob.k1, warm=warm, max_iter=300)
        if r["status"] != 0 and warm is not None:
            r = ipm_c.solve_packed(prob.N, x0, com, foot, gam, prob.mass, prob.k1, max_iter=300)
        res = R()
        res.status, res.iters, res.cost, res.viol = r["status"], r["iters"], r["cost"], r["viol"]
        X, U = r["X"].T, r["U"].T
        res.X = lambda n: X
        res.U = lambda n: U
        w = np.zeros(52 * prob.N + 20)
        for i in range(prob.N + 1):
            w[52 * i:52 * i + 20] = X[:, i]
        for i in range(prob.N):
            w[52 * i + 20:52 * i + 52] = U[:, i]
        res.w = w
        return res

    return RefMPC(initial, planner, params, com_ref, solver=solver, k1=k1, k2=0.1)


# (20, 0, 1100): 1100 ticks at the benchmark horizon from the standing start through nine steps and the whole push window 800 < t < 900
@pytest.mark.parametrize("N,t0,t1", [(10, 0, 60), (10, 180, 330), (20, 0, 1100)])
def test_closed_loop_com_within_1mm(pkg, N, t0, t1):
    from oracle.mpc_ref import surrogate_walk
    from oracle.walk import load_walk
    from cmpc_b200.centroidal_mpc_vertices import centroidal_mpc
    trajs = []
    for which in ("gpu", "oracle"):
        planner, com_ref, params, initial = load_walk()
        params["N"] = N
        if t0 > 0:
            initial["com"]["pos"] = np.array([com_ref["pos_x"][t0], com_ref["pos_y"][t0], 0.72])
            initial["com"]["vel"] = np.array([com_ref["vel_x"][t0], com_ref["vel_y"][t0], 0.0])
        mpc = centroidal_mpc(initial, planner, params, com_ref, None, None) if which == "gpu" else _oracle_mpc(initial, planner, params, com_ref)
        trajs.append(surrogate_walk(mpc, initial, t0, t1, params["mass"], hw_trace=initial["hw_meas"]))
        if which == "gpu":
            plan_gpu = copy.deepcopy([s["pos"] for s in planner.plan])
        else:
            plan_ref = [s["pos"] for s in planner.plan]
    a, b = trajs
    assert a.shape == b.shape == (t1 - t0, 12)
    assert np.abs(a[:, 0:3] - b[:, 0:3]).max() <= 1e-3, np.abs(a[:, 0:3] - b[:, 0:3]).max()          # CoM position: 1 mm
    assert np.abs(a[:, 3:6] - b[:, 3:6]).max() <= 1e-2
    assert np.abs(np.array(plan_gpu) - np.array(plan_ref)).max() <= 1e-3                               # step adjustment write-back (:669-672)


def test_drop_in_class_surface(pkg):
    """Constructor / solve() / model_state contract of the reference class (:7, :358-366, :480-683)."""
    from oracle.walk import load_walk
    from cmpc_b200.centroidal_mpc_vertices import centroidal_mpc
    from cmpc_b200.centroidal_mpc_vertices_payload import centroidal_mpc as payload_mpc
    planner, com_ref, params, initial = load_walk()
    mpc = centroidal_mpc(initial, planner, params, com_ref, None, None)
    cur = {"com": {"pos": np.array([0.0, 0.0, 0.72]), "vel": np.zeros(3)}, "hw": {"val": np.zeros(3)},
           "lfoot": {"pos": initial["lfoot"]["pos"]}, "rfoot": {"pos": initial["rfoot"]["pos"]}}
    ms, contact = mpc.solve(cur, 0)
    assert contact == "ds" and ms is mpc.model_state
    assert set(ms) == {"com", "hw", "theta_hat", "ang_contact_left", "pos_contact_left", "ang_contact_right",
                       "pos_contact_right", "mpc_new_contact", "counter"}
    assert ms["com"]["pos"].shape == (3,) and ms["com"]["acc"].shape == (3,) and ms["hw"]["dot"].shape == (3,)
    assert abs(ms["com"]["acc"][2]) < 0.05 and mpc.x.shape == (20,) and mpc.u.shape == (32,) and mpc.x_collect.shape == (20, params["N"] + 1)
    assert (mpc.k1, mpc.k2) == (4.0, 0.1)
    p2 = payload_mpc(initial, planner, params, com_ref, None, None)
    assert (p2.k1, p2.k2) == (7.0, 1.0)                                                                # payload file :27-31
    ms2, _ = p2.solve(cur, 0)
    assert np.isfinite(ms2["com"]["pos"]).all()
    with pytest.raises(IndexError):
        mpc.solve(cur, 1971)                                                                           # the reference's IndexError (:567)
    mpc.reset_update_swing_trj()


def test_payload_mass_sweep_and_long_horizon(pkg, walk_ticks):
    """BASELINE configs 4 and 5 in small: per-instance mass / k1 = 7 sweep, and horizon N = 60."""
    w = walk_ticks[20]
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 150, 128)                      # standing / early ticks stay feasible for heavier robots
    s = pkg.BatchSolver(20, 128, device=0)
    mass = float(w["mass"]) + rng.uniform(0, 10, 128)
    out = s.solve_host(w["x0"][idx], w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx], mass, 7.0, 0)
    conv = out["status"] == 0
    assert conv.mean() > 0.9 and out["viol"][conv].max() <= 1e-6
    fz = out["u0"][:, :24].reshape(-1, 8, 3)[:, :, 2].sum(axis=1)
    assert np.abs(fz[conv] / (mass[conv] * 9.81) - 1).max() < 0.2          # heavier robot -> proportionally larger vertical force
    # N = 60: refs re-sampled from the N = 20 tick inputs are not available, so stack three copies of the horizon of a
    # standing tick (constant references): exercises NMAX-sized tables and a 60-stage sweep
    k = 5
    com = np.tile(w["com_ref"][k], (3, 1)); foot = np.tile(w["foot_ref"][k], (3, 1)); gam = np.concatenate([np.tile(w["gamma"][k][:20], (3, 1)), w["gamma"][k][20:]])
    s60 = pkg.BatchSolver(60, 4, device=0)
    o60 = s60.solve_host(np.tile(w["x0"][k], (4, 1)), np.tile(com, (4, 1, 1)), np.tile(foot, (4, 1, 1)), np.tile(gam, (4, 1, 1)), float(w["mass"]), 4.0, 0)
    assert (o60["status"] == 0).all() and o60["viol"].max() <= 1e-6
