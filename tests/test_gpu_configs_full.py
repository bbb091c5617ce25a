"""BASELINE.json configs 3, 4 and 5 at one GPU's share of their full size (8192 perturbed states, 4096 payload masses,
2048 instances at horizon N = 60).  No oracle at these sizes: the checks are the size-independent ones (every instance
counted as converged satisfies all rows and dynamics to 1e-6; a warm re-solve from the converged iterate reproduces
the cost; the vertical contact force carries the weight).  A summary goes to gpurun_out/configs_full.json."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUMMARY = {}


def _record(name, **kw):
    SUMMARY[name] = kw
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "configs_full.json"), "w") as f:
            json.dump(SUMMARY, f, indent=1)


def test_config3_perturbed_states_8192(pkg, walk_ticks):
    N, B = 20, 8192
    w = walk_ticks[N]
    rng = np.random.default_rng(1)
    idx = rng.integers(0, len(w["x0"]), B)
    x0 = w["x0"][idx].copy()                                  # SURVEY.md 8d recipe
    x0[:, 0:3] += rng.normal(0, 0.01, (B, 3)); x0[:, 2] = np.minimum(x0[:, 2], 0.759)
    x0[:, 3:6] += rng.normal(0, 0.05, (B, 3))
    x0[:, 6:9] = rng.normal(0, 1.0, (B, 3)) * np.array([0.88, 0.63, 0.20])
    x0[:, 9:12] = rng.normal(0, 2.0, (B, 3))
    args = (x0, w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx], float(w["mass"]), float(w["k1"]))
    s = pkg.BatchSolver(N, B, device=0)
    out = s.solve_host(*args, 0)
    st = s.last_stats()
    conv = out["status"] == 0
    assert conv.mean() > 0.5
    assert out["viol"][conv].max() <= 1e-6
    again = s.solve_host(*args, 2)                            # full warm start from the converged iterate: a fixed point
    both = conv & (again["status"] == 0)
    assert both.sum() >= 0.99 * conv.sum()
    rel = np.abs(again["cost"][both] - out["cost"][both]) / np.maximum(1.0, np.abs(out["cost"][both]))
    assert np.quantile(rel, 0.99) <= 1e-6                     # (a handful may move to another KKT point of the non-convex NLP)
    _record("config3_perturbed", batch=B, horizon=N, converged_fraction=float(conv.mean()), cold_solves_per_s=float(conv.sum() / st["kernel_ms"] * 1e3),
            kernel_ms=st["kernel_ms"], iters_per_solve=st["iters"] / B, status_hist=np.bincount(out["status"], minlength=6).tolist(),
            max_viol_converged=float(out["viol"][conv].max()), resolve_rel_cost_q99=float(np.quantile(rel, 0.99)))


def test_config4_payload_masses_4096(pkg, walk_ticks):
    N, B = 20, 4096
    w = walk_ticks[N]
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 150, B)                             # standing / early ticks stay feasible for heavier robots
    mass = float(w["mass"]) + rng.uniform(0, 10, B)
    s = pkg.BatchSolver(N, B, device=0)
    out = s.solve_host(w["x0"][idx], w["com_ref"][idx], w["foot_ref"][idx], w["gamma"][idx], mass, 7.0, 0)   # k1 = 7: payload file :27-31
    st = s.last_stats()
    conv = out["status"] == 0
    assert conv.mean() > 0.9 and out["viol"][conv].max() <= 1e-6
    fz = out["u0"][:, :24].reshape(-1, 8, 3)[:, :, 2].sum(axis=1)
    ratio = fz[conv] / (mass[conv] * 9.81)
    assert np.abs(ratio - 1).max() < 0.2
    _record("config4_payload", batch=B, horizon=N, k1=7.0, converged_fraction=float(conv.mean()), cold_solves_per_s=float(conv.sum() / st["kernel_ms"] * 1e3),
            kernel_ms=st["kernel_ms"], max_viol_converged=float(out["viol"][conv].max()), fz_over_weight=[float(ratio.min()), float(ratio.max())])


def test_config5_long_horizon_2048(pkg, walk_ticks):
    from oracle.walk import load_walk                          # fixture loader (planner tables of the recorded walk)
    from cmpc_b200.assembly import PlanTables, ReferenceTables, assemble_tick, pack_instances
    N, B = 60, 2048
    planner, com_ref, params, initial = load_walk()
    params = dict(params, N=N)
    tables, refs = PlanTables(planner.plan), ReferenceTables(com_ref, planner)
    w = walk_ticks[20]
    rng = np.random.default_rng(3)
    ticks = rng.integers(0, 1900, B)                           # t + 60 stays inside the 1971-row reference tables

    def instance(t):
        x = w["x0"][t]                                         # state the recorded (N = 20) walk had at tick t
        cur = {"com": {"pos": x[0:3], "vel": x[3:6]}, "hw": {"val": x[6:9]}, "lfoot": {"pos": [0, 0, x[12]]}, "rfoot": {"pos": [0, 0, x[16]]}}
        return assemble_tick(tables, refs, planner.plan, params, cur, x[9:12], int(t))

    now = pack_instances([instance(t) for t in ticks])
    nxt = pack_instances([instance(t + 1) for t in ticks])
    s = pkg.BatchSolver(N, B, device=0)
    out = s.solve_host(*now, float(w["mass"]), float(w["k1"]), 0)
    st0 = s.last_stats()
    conv = out["status"] == 0
    assert conv.mean() > 0.95 and out["viol"][conv].max() <= 1e-6
    out1 = s.solve_host(*nxt, float(w["mass"]), float(w["k1"]), 4)          # the next tick, warm-started on the device (automatic shift)
    st1 = s.last_stats()
    conv1 = out1["status"] == 0
    assert conv1.mean() > 0.95 and out1["viol"][conv1].max() <= 1e-6
    fp = s.footprint()
    _record("config5_long_horizon", batch=B, horizon=N, converged_fraction_cold=float(conv.mean()), converged_fraction_warm=float(conv1.mean()),
            cold_solves_per_s=float(conv.sum() / st0["kernel_ms"] * 1e3), warm_solves_per_s=float(conv1.sum() / st1["kernel_ms"] * 1e3),
            iters_per_solve_cold=st0["iters"] / B, iters_per_solve_warm=st1["iters"] / B, iterate_kb_per_instance=fp["iterate_bytes_per_instance"] / 1e3, scratch_mb_per_slot=fp["scratch_bytes_per_slot"] / 1e6, slots=fp["slots"],
            max_viol_converged=float(max(out["viol"][conv].max(), out1["viol"][conv1].max())))
