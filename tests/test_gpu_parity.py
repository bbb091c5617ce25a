"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle's golden vectors
(tests/golden/golden_N*.npz, produced by tests/golden/make_golden.py with oracle/ipm_py.py) and against
size-independent properties on full-size batches."""
import numpy as np
import pytest

from parity import COST_TOL, U0_TOL, VIOL_TOL, X1_TOL, cost_err, golden_errors, u0_err

pytestmark = pytest.mark.gpu


def _solve_golden(pkg, g, N, warm=0, **over):
    B = len(g["ticks"])
    s = pkg.BatchSolver(N, B, device=0, **over)
    out = s.solve_host(g["x0"], g["com_ref"], g["foot_ref"], g["gamma"], float(g["mass"]), float(g["k1"]), warm)
    return s, out


@pytest.mark.parametrize("N", [10, 20])
def test_golden_parity_cold(pkg, golden, N):
    g = golden[N]
    s, out = _solve_golden(pkg, g, N)
    ok = g["status"] == 0
    assert ok.sum() >= len(ok) - 1
    assert (out["status"][ok] == 0).all(), out["status"]
    assert out["viol"][ok].max() <= VIOL_TOL
    for k in np.flatnonzero(ok):
        ec, ex, eu = golden_errors(g, k, out["cost"][k], out["x1"][k], out["u0"][k])
        assert ec <= COST_TOL and ex <= X1_TOL and eu <= U0_TOL, (int(g["ticks"][k]), ec, ex, eu)
    X, U = s.trajectory(len(ok))
    assert np.abs(X[:, 1] - out["x1"]).max() == 0.0 and np.abs(U[:, 0] - out["u0"]).max() == 0.0


def test_k2_independence_and_payload_gain(pkg, golden):
    """k2 cancels out of the NLP: the ABI has no k2 at all; k1 is per instance (payload k1 = 7)."""
    g = golden[10]
    B = len(g["ticks"])
    s = pkg.BatchSolver(10, 2 * B, device=0)
    k1 = np.concatenate([np.full(B, 4.0), np.full(B, 7.0)])
    cat = lambda a: np.concatenate([a, a])
    out = s.solve_host(cat(g["x0"]), cat(g["com_ref"]), cat(g["foot_ref"]), cat(g["gamma"]), float(g["mass"]), k1, 0)
    ok = g["status"] == 0
    assert cost_err(out["cost"][:B], g["cost"])[ok].max() <= COST_TOL
    conv = (out["status"][B:] == 0)
    assert conv.mean() > 0.7
    assert np.abs(out["cost"][B:][conv] - out["cost"][:B][conv]).max() > 1e-9      # a different problem indeed
    assert out["viol"][B:][conv].max() <= VIOL_TOL


def test_warm_start_modes_agree(pkg, golden):
    g = golden[10]
    B = len(g["ticks"])
    s, cold = _solve_golden(pkg, g, 10)
    args = (g["x0"], g["com_ref"], g["foot_ref"], g["gamma"], float(g["mass"]), float(g["k1"]))
    full = s.solve_host(*args, 2)
    prim = s.solve_host(*args, 1)
    ok = (cold["status"] == 0)
    for o in (full, prim):
        assert (o["status"][ok] == 0).all()
        assert cost_err(o["cost"], cold["cost"])[ok].max() <= COST_TOL
        assert np.abs(o["x1"][:, :12] - cold["x1"][:, :12])[ok].max() <= X1_TOL
    # re-solving from the converged point is much cheaper than a cold solve
    assert full["iters"][ok].mean() < 0.5 * cold["iters"][ok].mean()


def test_shifted_and_automatic_warm_start_on_consecutive_ticks(pkg, walk_ticks):
    """WARM_SHIFTED (3) and WARM_AUTO (4) on consecutive ticks of the recorded walk: same KKT points as a cold solve of
    the same tick (cost 1e-6) for every tick that converges both ways, and cheaper than cold on average."""
    N = 20
    w = walk_ticks[N]
    ticks = np.array([150, 199, 230, 255, 262, 268, 271, 300, 640, 805, 1455, 1700])
    B = len(ticks)
    args = lambda tt: (w["x0"][tt], w["com_ref"][tt], w["foot_ref"][tt], w["gamma"][tt], float(w["mass"]), float(w["k1"]))
    cold = pkg.BatchSolver(N, B, device=0).solve_host(*args(ticks), 0)
    assert (cold["status"] == 0).all()
    for mode in (3, 4):
        s = pkg.BatchSolver(N, B, device=0)
        s.solve_host(*args(ticks - 1), 0)
        o = s.solve_host(*args(ticks), mode)
        assert (o["status"] == 0).all(), (mode, o["status"])
        assert o["viol"].max() <= VIOL_TOL
        same = cost_err(o["cost"], cold["cost"]) <= COST_TOL
        assert same.sum() >= B - 2, (mode, cost_err(o["cost"], cold["cost"]))          # (a non-convex NLP: a tick may have a second KKT point)
        assert cost_err(o["cost"], cold["cost"]).max() <= 1e-4
        assert o["iters"].mean() < cold["iters"].mean()


def test_reset_warm_mask_and_retry_accounting(pkg, golden):
    """cmpc_reset_warm with a per-instance mask: flagged instances start cold on the next warm solve (same iteration
    count as a cold solve), the others keep their warm start."""
    g = golden[10]
    B = len(g["ticks"])
    args = (g["x0"], g["com_ref"], g["foot_ref"], g["gamma"], float(g["mass"]), float(g["k1"]))
    s = pkg.BatchSolver(10, B, device=0)
    cold = s.solve_host(*args, 0)
    warm = s.solve_host(*args, 2)
    mask = np.zeros(B, bool); mask[::2] = True
    s.reset_warm(mask)
    mixed = s.solve_host(*args, 2)
    ok = cold["status"] == 0
    assert np.array_equal(mixed["iters"][mask & ok], cold["iters"][mask & ok])          # cold again, bit for bit the same solve
    assert np.array_equal(mixed["cost"][mask & ok], cold["cost"][mask & ok])
    assert mixed["iters"][~mask & ok].mean() < 0.75 * cold["iters"][~mask & ok].mean()      # the others kept their warm start
    assert warm["iters"][ok].mean() < 0.6 * cold["iters"][ok].mean()
    s.reset_warm()
    again = s.solve_host(*args, 2)
    assert np.array_equal(again["iters"][ok], cold["iters"][ok])


def test_operations_on_different_streams_are_ordered(pkg, golden):
    """A solve on a user stream followed by handle-stream operations (trajectory export, snapshot) and back: the handle
    orders them (no explicit synchronisation by the caller)."""
    import torch
    g = golden[10]
    B = len(g["ticks"])
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float64), device=dev)
    s = pkg.BatchSolver(10, B, device=0)
    ins = (t(g["x0"]), t(g["com_ref"]), t(g["foot_ref"]), t(g["gamma"]), t(np.full(B, float(g["mass"]))), t(np.full(B, float(g["k1"]))))
    user = torch.cuda.Stream(dev)
    with torch.cuda.stream(user):
        out = s.solve_device(*ins, 0, stream=user.cuda_stream)
    X, U = s.trajectory(B)                                   # handle's own stream: must see the finished solve
    torch.cuda.synchronize()
    assert np.abs(X[:, 1] - out["x1"].cpu().numpy()).max() == 0.0 and np.abs(U[:, 0] - out["u0"].cpu().numpy()).max() == 0.0
    s.warm_save(B)
    with torch.cuda.stream(user):
        out2 = s.solve_device(*ins, 2, stream=user.cuda_stream)
    s.warm_restore(B)                                        # handle's stream, ordered after the solve on `user`
    with torch.cuda.stream(user):
        out3 = s.solve_device(*ins, 2, stream=user.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out2["iters"].cpu().numpy(), out3["iters"].cpu().numpy())     # same snapshot, same warm solve
    assert np.abs(out2["cost"].cpu().numpy() - out3["cost"].cpu().numpy()).max() == 0.0


def test_device_entry_point_matches_host_entry_point(pkg, golden):
    import torch
    g = golden[10]
    B = len(g["ticks"])
    s, ref = _solve_golden(pkg, g, 10)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float64), device=dev)
    s2 = pkg.BatchSolver(10, B, device=0)
    out = s2.solve_device(t(g["x0"]), t(g["com_ref"]), t(g["foot_ref"]), t(g["gamma"]),
                          t(np.full(B, float(g["mass"]))), t(np.full(B, float(g["k1"]))), 0)
    torch.cuda.synchronize()
    assert np.array_equal(out["status"].cpu().numpy(), ref["status"])
    assert np.abs(out["cost"].cpu().numpy() - ref["cost"]).max() == 0.0          # same kernel, same bits
    assert np.abs(out["u0"].cpu().numpy() - ref["u0"]).max() == 0.0


def test_batch_order_and_replication_invariance(pkg, walk_ticks):
    """Instances are independent: permuting / replicating the batch permutes / replicates the results bit for bit."""
    w = walk_ticks[10]
    rng = np.random.default_rng(0)
    idx = rng.integers(0, len(w["x0"]), 96)
    s = pkg.BatchSolver(10, 192, device=0)
    a = s.solve_host(w["x0"][idx], w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx], float(w["mass"]), float(w["k1"]), 0)
    perm = rng.permutation(96)
    idx2 = np.concatenate([idx[perm], idx])
    b = s.solve_host(w["x0"][idx2], w["com_ref"][idx2], w["foot_ref"][idx2], w["gamma"][idx2], float(w["mass"]), float(w["k1"]), 0)
    assert np.array_equal(b["cost"][:96], a["cost"][perm]) and np.array_equal(b["cost"][96:], a["cost"])
    assert np.array_equal(b["u0"][96:], a["u0"]) and np.array_equal(b["iters"][:96], a["iters"][perm])


@pytest.mark.parametrize("N", [20])
def test_full_size_batch_properties(pkg, walk_ticks, N):
    """Batch 4096 at N = 20 (BASELINE config 2): every recorded tick converges, violations <= 1e-6, the
    angular-momentum row (:224) and friction pyramids hold on u0/x1, and the result equals the small-batch result."""
    w = walk_ticks[N]
    rng = np.random.default_rng(0)
    idx = rng.integers(0, len(w["x0"]), 4096)
    s = pkg.BatchSolver(N, 4096, device=0)
    out = s.solve_host(w["x0"][idx], w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx], float(w["mass"]), float(w["k1"]), 0)
    conv = out["status"] == 0
    assert conv.mean() >= 0.995, np.bincount(out["status"])      # cold start from the neutral guess; non-converged instances are reported, never counted
    assert out["viol"][conv].max() <= VIOL_TOL
    h0 = np.linalg.norm(w["x0"][idx][:, 6:9], axis=1)
    h1 = np.linalg.norm(out["x1"][:, 6:9], axis=1)
    assert (h1[conv] ** 2 <= h0[conv] ** 2 + 1e-6).all()
    f = out["u0"][:, :24].reshape(-1, 8, 3)
    gam0 = np.repeat(w["gamma"][idx][:, 0, :], 4, axis=1)
    on = (gam0 > 0.5) & conv[:, None]
    assert (np.abs(f[..., 0])[on] <= 0.5 * f[..., 2][on] + 1e-6).all() and (np.abs(f[..., 1])[on] <= 0.5 * f[..., 2][on] + 1e-6).all()
    s2 = pkg.BatchSolver(N, 64, device=0)
    small = s2.solve_host(w["x0"][idx[:64]], w["com_ref"][idx[:64]], w["foot_ref"][idx[:64]], w["gamma"][idx[:64]],
                          float(w["mass"]), float(w["k1"]), 0)
    assert np.array_equal(small["cost"], out["cost"][:64])


def test_infeasible_x0_is_reported_not_counted(pkg, golden):
    g = golden[10]
    x0 = g["x0"].copy()
    x0[:, 2] = 0.80                                      # CoM above the 0.76 bound of :230 at stage 0
    s = pkg.BatchSolver(10, len(x0), device=0)
    out = s.solve_host(x0, g["com_ref"], g["foot_ref"], g["gamma"], float(g["mass"]), float(g["k1"]), 0)
    assert (out["status"] != 0).all()


def test_batch_capacity_and_empty(pkg, golden):
    g = golden[10]
    s = pkg.BatchSolver(10, 2, device=0)
    with pytest.raises(pkg.CmpcError):
        s.solve_host(g["x0"][:3], g["com_ref"][:3], g["foot_ref"][:3], g["gamma"][:3], 40.0, 4.0, 0)
    one = s.solve_host(g["x0"][:1], g["com_ref"][:1], g["foot_ref"][:1], g["gamma"][:1], float(g["mass"]), float(g["k1"]), 0)
    assert one["status"][0] == 0


def test_solves_are_deterministic_and_independent_of_batch_order(pkg, walk_ticks):
    """The same instance gives bit-identical results run after run, warm or cold, wherever it sits in the batch and in
    whatever order the CTAs are launched (a data race or an uninitialised read in the kernel would show up here;
    compute-sanitizer is not available on the GPU pool)."""
    N, B = 20, 600                                            # more instances than the 592 resident CTA slots
    w = walk_ticks[N]
    rng = np.random.default_rng(5)
    idx = rng.integers(1, len(w["x0"]), B)
    arg = lambda ii: (w["x0"][ii], w["com_ref"][ii], w["foot_ref"][ii], w["gamma"][ii], float(w["mass"]), float(w["k1"]))
    keys = ("x1", "u0", "xN", "cost", "viol", "status", "iters")

    def run(order):
        s = pkg.BatchSolver(N, B, device=0)
        cold = s.solve_host(*arg(idx[order] - 1), 0)
        warm = s.solve_host(*arg(idx[order]), 2)                 # launch order = work of the cold solves
        s.close()
        inv = np.argsort(order)
        return [{k: np.array(o[k])[inv] for k in keys} for o in (cold, warm)]

    ident = np.arange(B)
    a, b, c = run(ident), run(ident), run(rng.permutation(B))
    for first, other in ((a, b), (a, c)):
        for o1, o2 in zip(first, other):
            for k in keys:
                assert np.array_equal(o1[k], o2[k]), k
