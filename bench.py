#!/usr/bin/env python
"""bench.py -- converged centroidal-MPC solves/s on recorded-walk batches (BASELINE.json configs).

A "step" is one MPC tick for a batch of independent instances.

  --config 2 (default, the metric's configuration, BASELINE.json configs[1]): batch 4096 per GPU, horizon N = 20,
      instances = ticks of the recorded surrogate walk sampled with replacement (seed = rank), each warm-started from
      the solution of ITS previous tick, which is resident in the solver handle on the device (the solver's normal
      operating mode: states, inputs, costates, slacks and multipliers stay in HBM across ticks).  Every timed step
      first restores that previous-tick state from a device snapshot (device-to-device copy, inside the timed region)
      and then solves the tick.  Extra keys: `cold_start` (same batch from the solver's own initial guess) and
      `rolling_replay` (consecutive ticks t, t+1, ... of every instance, each warm-started from the tick before).
  --config 3: disturbance-robustness sweep, 65 536 perturbed initial states (SURVEY.md 8d recipe) sharded over the
      GPUs (8192 per GPU when run on fewer than 8), horizon 20, cold start.
  --config 4: payload variant, k1 = 7, per-instance mass 40.05 + U(0, 10) kg, batch 4096 per GPU, cold start.
  --config 5: long horizon N = 60, batch 16 384 in total (sharded), warm replay as config 2.

  value   converged solves/s with inputs resident in HBM (cmpc_solve_device), CUDA events, max over ranks
  e2e     same metric through the host-buffer C-ABI call the drop-in class uses (cmpc_solve_host): pinned
          staging + H2D + kernels + D2H inside the timed region
  roofline  FP64: dense-convention flops (121,749 N per Riccati factorisation + 10,368 N per
          solve) / kernel time vs the DFMA peak measured live on this GPU; HBM: algorithmic bytes
          8 (123 N + 62) per solve vs MEASURED_PEAKS.json hbm_gbs; `traffic` = DRAM bytes of the solve kernel measured
          by ncu on this very shape (profiles/r02_dram_bytes.json, written by scripts/dram_bytes.sh), else null
  cpu_baseline  the restated reference (oracle/ipm_c.c, "oracle-R": IPOPT's termination test at the reference's
          tol = 1e-3, primal warm start from the previous tick as :630-631) on the host cores, one process per core,
          wall clock, bounded sample of the same workload; the tight-tolerance oracle-T rate is reported beside it

`--impl reference` times only that restated CPU path (rank 0), same JSON contract.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged centroidal-MPC solves/sec"
UNIT = "solves/s"
F_FACT, F_SOLVE = 121749.0, 10368.0           # dense-convention flops per stage
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------------------------- workloads
def load_ticks(N):
    return np.load(os.path.join(GOLDEN, "walk_ticks_N%d.npz" % N))


def walk_tables():
    """Planner / reference tables of the recorded walk (tests/golden/walk_inputs.npz, made by importing the reference's
    own planner) in the shapes the reference's `centroidal_mpc` constructor takes."""
    d = np.load(os.path.join(GOLDEN, "walk_inputs.npz"))

    class Planner:
        pass

    pl = Planner()
    pl.plan = [{"pos": np.array(d["plan_pos"][j], float), "ang": np.array(d["plan_ang"][j], float), "ss_duration": int(d["plan_ss"][j]),
                "ds_duration": int(d["plan_ds"][j]), "foot_id": "lfoot" if int(d["plan_foot"][j]) == 0 else "rfoot"} for j in range(len(d["plan_ss"]))]
    pl.position_contacts_ref = {"contact_left": np.array(d["contact_left"], float), "contact_right": np.array(d["contact_right"], float)}
    com_ref = {k[4:]: np.array(d[k], float) for k in d.files if k.startswith("ref_")}
    params = {"g": 9.81, "h": 0.72, "foot_size": 0.1, "world_time_step": 0.01, "first_swing": "rfoot", "N": 10, "mass": float(d["mass"]),
              "update_contact": "YES", "mpc_rate": 1}
    initial = {"lfoot": {"pos": np.array(d["lfoot0"], float)}, "rfoot": {"pos": np.array(d["rfoot0"], float)},
               "com": {"pos": np.array([0.0, 0.0, 0.72]), "vel": np.zeros(3)}, "hw": {"val": np.zeros(3)}, "hw_meas": np.array(d["hw_meas"], float)}
    return pl, com_ref, params, initial


def replay_workload(N, batch, seed, back=2, ahead=0, headroom=None):
    """Instances = ticks of the recorded walk sampled with replacement; returns the inputs of ticks t-back .. t+ahead.
    `headroom` (>= ahead) fixes the range the ticks are drawn from, so the sample does not depend on `ahead`."""
    headroom = ahead if headroom is None else max(headroom, ahead)
    rng = np.random.default_rng(seed)
    if N in (10, 20):
        w = load_ticks(N)
        idx = rng.integers(back, len(w["x0"]) - headroom, batch)
        take = lambda ii: tuple(np.ascontiguousarray(w[k][ii]) for k in ("x0", "com_ref", "foot_ref", "gamma"))
        return [take(idx + o) for o in range(-back, ahead + 1)], float(w["mass"]), float(w["k1"]), idx
    # other horizons: re-assemble from the walk's tables at the states the recorded N = 20 walk had
    import cmpc_loader
    cmpc_loader.load()
    from cmpc_b200.assembly import PlanTables, ReferenceTables, assemble_tick, pack_instances
    planner, com_ref, params, _ = walk_tables()
    params = dict(params, N=N)
    tables, refs = PlanTables(planner.plan), ReferenceTables(com_ref, planner)
    w = load_ticks(20)
    idx = rng.integers(back, 1970 - N - headroom - 1, batch)

    def instance(t):
        x = w["x0"][t]
        cur = {"com": {"pos": x[0:3], "vel": x[3:6]}, "hw": {"val": x[6:9]}, "lfoot": {"pos": [0, 0, x[12]]}, "rfoot": {"pos": [0, 0, x[16]]}}
        return assemble_tick(tables, refs, planner.plan, params, cur, x[9:12], int(t))

    uniq = {}
    def tick(t):
        if t not in uniq:
            uniq[t] = instance(t)
        return uniq[t]
    return [pack_instances([tick(int(t) + o) for t in idx]) for o in range(-back, ahead + 1)], float(w["mass"]), float(w["k1"]), idx


def perturbed_workload(batch, seed):
    """BASELINE config 3: base ticks of the N = 20 walk, perturbed CoM / momentum / theta_hat."""
    w = load_ticks(20)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(w["x0"]), batch)
    x0 = w["x0"][idx].copy()
    x0[:, 0:3] += rng.normal(0, 0.01, (batch, 3)); x0[:, 2] = np.minimum(x0[:, 2], 0.759)
    x0[:, 3:6] += rng.normal(0, 0.05, (batch, 3))
    x0[:, 6:9] = rng.normal(0, 1.0, (batch, 3)) * np.array([0.88, 0.63, 0.20])
    x0[:, 9:12] = rng.normal(0, 2.0, (batch, 3))
    return (x0, np.ascontiguousarray(w["com_ref"][idx]), np.ascontiguousarray(w["foot_ref"][idx]), np.ascontiguousarray(w["gamma"][idx])), float(w["mass"]), float(w["k1"])


def payload_workload(batch, seed):
    """BASELINE config 4 (i): payload variant k1 = 7 (payload file :27-31), mass = 40.05 + U(0, 10) kg, standing / early ticks."""
    w = load_ticks(20)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 150, batch)
    mass = float(w["mass"]) + rng.uniform(0, 10, batch)
    return tuple(np.ascontiguousarray(w[k][idx]) for k in ("x0", "com_ref", "foot_ref", "gamma")), mass, 7.0


# ----------------------------------------------------------------------------------------------- CPU baseline
def _oracle_prev(args):
    from oracle import ipm_c as orc
    N, prev, mass, k1, opts = args
    r0 = orc.solve_packed(N, *prev, mass, k1, **opts)
    return (r0["X"], r0["U"]) if r0["status"] == 0 else None


def _oracle_cur(args):
    """One tick solved the way the reference does it: warm-started from the previous tick's primal solution
    (`opt.set_initial`, MPC file :630-631; slacks / multipliers / barrier restart)."""
    from oracle import ipm_c as orc
    N, cur, mass, k1, warm, opts = args
    t0 = time.perf_counter()
    r = orc.solve_packed(N, *cur, mass, k1, warm=warm, **opts)
    if r["status"] != 0 and warm is not None:
        r = orc.solve_packed(N, *cur, mass, k1, **opts)
    return {"status": r["status"], "iters": r["iters"], "secs": time.perf_counter() - t0}


def cpu_reference_rate(N, n_instances, seed=0, procs=None, tight=False):
    """Restated reference (oracle/ipm_c.c) on the host cores, one single-threaded process per core.  Phase 1 (untimed): tick
    t-1 of every instance, for the warm starts.  Phase 2 (timed, WALL clock around the whole pool): tick t."""
    from multiprocessing import get_context
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import ipm_c
    ipm_c.build()
    opts = {} if tight else dict(ipm_c.ORACLE_R)
    procs = procs or os.cpu_count()
    (prev, cur), mass, k1, idx = replay_workload(N, n_instances, seed, back=1)
    with get_context("fork").Pool(procs) as pool:
        warms = pool.map(_oracle_prev, [(N, tuple(a[b] for a in prev), mass, k1, opts) for b in range(n_instances)], chunksize=4)
        jobs = [(N, tuple(a[b] for a in cur), mass, k1, warms[b], opts) for b in range(n_instances)]
        t0 = time.perf_counter()
        res = pool.map(_oracle_cur, jobs, chunksize=4)
        wall = time.perf_counter() - t0
    conv = sum(1 for r in res if r["status"] == 0)
    return {"rate": conv / wall, "wall_s": wall, "converged": conv, "n": n_instances, "procs": procs, "iters": float(np.mean([r["iters"] for r in res])),
            "ms_per_solve_per_core": 1e3 * float(np.mean([r["secs"] for r in res]))}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows, self.gpu, self.p = [], gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        mhz = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(mhz)}


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def config_of(a, world):
    """(N, batch per GPU, description) of the selected BASELINE config."""
    if a.config == 3:
        B = a.batch or (65536 // world if world >= 8 else 8192)
        return 20, B, "config 3: disturbance-robustness sweep, %d perturbed initial CoM / momentum states per GPU (65 536 over 8 GPUs), horizon N=20, cold start" % B
    if a.config == 4:
        B = a.batch or 4096
        return 20, B, "config 4: payload variant k1=7, per-instance mass 40.05 + U(0,10) kg, batch %d per GPU, horizon N=20, cold start" % B
    if a.config == 5:
        B = a.batch or max(16384 // world, 1)
        return 60, B, "config 5: long horizon N=60, batch 16 384 in total = %d per GPU, warm replay of recorded-walk ticks" % B
    N = a.horizon
    B = a.batch or 4096
    return N, B, ("replay of recorded surrogate-walk ticks, batch %d per GPU, horizon N=%d, warm start from the previous tick's "
                  "device-resident solution" % (B, N))


def main():
    # stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner, torchrun
    # notices of child processes) is sent to stderr for the duration of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config (2 = the metric's configuration)")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0 = the config's own size)")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU baseline sample (0 = 128 per core, about 10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip cold-start / rolling-replay / latency / fleet extras (A/B runs, ncu captures)")
    ap.add_argument("--samples", type=int, default=4, help="independent instance samples the timed steps cycle through (configs 2 and 5): a step is a makespan and moves by +-10 %% with the sampled ticks")
    ap.add_argument("--warm-mode", type=int, default=4, help="warm-start mode of the replay (4 = automatic shift, 2 = full, 3 = shifted)")
    ap.add_argument("--cfg", action="append", default=[], help="solver option override key=value (experiments only; the default run uses the library defaults)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    K, W = a.steps, max(a.warmup, 0)
    N, B, workload = config_of(a, world)
    warm_replay = a.config in (2, 5)
    WM = a.warm_mode
    config = {"workload": workload, "baseline_config": a.config, "batch_per_gpu": B, "horizon": N,
              "warm_start": ("mode %d (4 = per instance: shifted while a landing is inside the horizon, else full), device snapshot of tick t-1 restored every step" % WM)
                            if warm_replay else "cold (solver's own initial guess)",
              "launches_per_step": "cmpc_order_kernel (launch order from tick t-1's work; warm only) + cmpc_solve_kernel (persistent CTAs, retries of failed instances re-enter its work queue); two small memsets",
              "cache": "%s", "parallelism": "instances sharded over %d GPU(s), no collective on the hot path" % world,
              "makespan_note": "a step is the makespan of the batch on the resident CTA slots: it moves by about +-10 % with the sampled ticks / any change of the rounding (which instance is the straggler); DESIGN.md section 5"}

    # ------------------------------------------------------------------ reference arm: restated CPU path only
    if a.impl == "reference":
        if rank != 0:
            return
        ncpu = os.cpu_count()
        Nr = N if N in (10, 20) else 20
        per_step = a.cpu_sample or 32 * ncpu
        runs = []
        for s in range(W + K):
            r = cpu_reference_rate(Nr, per_step, seed=s)
            if s >= W:
                runs.append(r)
        value = sum(r["converged"] for r in runs) / sum(r["wall_s"] for r in runs)
        config["cache"] = "n/a (CPU)"
        config["same_config_note"] = ("each step = a bounded sample of %d instances of the config-2 tick replay at N=%d (not %d per step): a throughput "
                                      "estimate of the same per-instance work" % (per_step, Nr, B))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * np.mean([r["wall_s"] for r in runs]), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (recorded surrogate walk)", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": runs[0]["procs"], "kind": "port",
                                 "sample": "%d instances per step of the N=%d tick replay, primal warm start from tick t-1, restated reference (oracle/ipm_c.c with the "
                                           "reference's IPOPT settings tol=1e-3 / constr_viol_tol=1e-4: CasADi/IPOPT not installable offline), gcc -O3 -march=native, one "
                                           "process per core, wall clock, %.1f iterations/solve" % (per_step, Nr, runs[-1]["iters"])},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    import torch
    import cmpc_loader
    pkg = cmpc_loader.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
    over = {k: float(v) for k, v in (kv.split("=") for kv in a.cfg)}
    for k in ("max_iter", "ls_max", "stall_window", "stall_final", "jam_window"):
        if k in over:
            over[k] = int(over[k])
    R = 0 if (a.no_extras or a.config != 2) else min(max(K, 1), 32)  # ticks of the rolling replay
    NS = max(1, a.samples) if warm_replay else 1                     # independent instance samples the steps cycle through
    stream = torch.cuda.Stream(dev)            # the solver's launches and the timing events share this stream
    samples = []
    for j in range(NS):
        smp = {}
        if warm_replay:
            # (the sampled ticks do not depend on the flags: always 32 ticks of headroom, of which the rolling replay uses R)
            ticks, mass, k1, idx = replay_workload(N, B, seed=rank * NS + j, back=2, ahead=R if j == 0 else 0, headroom=32)
            prev2, prev, cur = ticks[0], ticks[1], ticks[2]
            mass_h, k1_h = np.full(B, mass), np.full(B, k1)
            if j == 0:
                ticks0 = ticks
        elif a.config == 3:
            cur, mass, k1 = perturbed_workload(B, seed=1000 + rank)
            mass_h, k1_h = np.full(B, mass), np.full(B, k1)
        else:
            cur, mass_h, k1 = payload_workload(B, seed=2000 + rank)
            k1_h = np.full(B, k1)
        sv = pkg.BatchSolver(N, B, device=local, **over)
        mass_t, k1_t = t(mass_h), t(k1_h)
        cur_t = [t(x) for x in cur]
        if warm_replay:
            # the loop's steady state, untimed: tick t-2 from cold, tick t-1 warm-started from it.  The snapshot then holds what a
            # running controller has on the device when tick t arrives: the iterate of t-1 and the work its (warm) solve took,
            # which orders the launch of tick t (longest expected first)
            sv.solve_device(*[t(x) for x in prev2], mass_t, k1_t, 0)
            out = sv.solve_device(*[t(x) for x in prev], mass_t, k1_t, WM)
            torch.cuda.synchronize()
            sv.warm_save(B)
        else:
            out = sv.solve_device(*cur_t, mass_t, k1_t, 0)
            torch.cuda.synchronize()
        smp.update(solver=sv, cur=cur, cur_t=cur_t, mass_h=mass_h, k1_h=k1_h, mass_t=mass_t, k1_t=k1_t, out=out)
        samples.append(smp)
    solver, cur, cur_t, mass_h, k1_h, mass_t, k1_t, out = (samples[0][k] for k in ("solver", "cur", "cur_t", "mass_h", "k1_h", "mass_t", "k1_t", "out"))
    ticks = ticks0 if warm_replay else None
    fp = solver.footprint()
    config["cache"] = ("iterates %.2f GB per GPU (33 KB per instance at N=20) + scratch of the %d resident CTA slots %.2f GB, streamed every iteration "
                       "(> 126 MB L2); no extra flush" % (B * fp["iterate_bytes_per_instance"] / 1e9, min(fp["slots"], B),
                                                          min(fp["slots"], B) * fp["scratch_bytes_per_slot"] / 1e9))
    config["samples"] = "%d independent instance samples (seeds %d..%d), the timed steps cycle through them" % (NS, rank * NS, rank * NS + NS - 1)
    torch.cuda.set_stream(stream)
    step_no = [0]

    def step_device(j=None):
        if j is None:
            j = step_no[0] % NS
            step_no[0] += 1
        m = samples[j]
        if warm_replay:
            m["solver"].warm_restore(B, stream.cuda_stream)
            m["solver"].solve_device(*m["cur_t"], m["mass_t"], m["k1_t"], WM, out=m["out"], stream=stream.cuda_stream)
        else:
            m["solver"].solve_device(*m["cur_t"], m["mass_t"], m["k1_t"], 0, out=m["out"], stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]      # per-step marks inside the one timed region
    step_no[0] = 0
    e0.record(stream)
    marks[0].record(stream)
    for k in range(K):
        step_device()
        marks[k + 1].record(stream)
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    step_ms = [marks[k].elapsed_time(marks[k + 1]) for k in range(K)]
    clocks = sampler.stop() if rank == 0 else None
    # per sample: converged count, iteration statistics and the kernels' duration of one step (CUDA events on the launching stream)
    per = []
    for j in range(NS):
        step_device(j)
        torch.cuda.synchronize()
        stj = samples[j]["solver"].last_stats()
        stat = samples[j]["out"]["status"].cpu().numpy()
        per.append({"conv": int((stat == 0).sum()), "nfact": stj["nfact"], "iters": stj["iters"], "kernel_ms": stj["kernel_ms"], "launches": stj["launches"], "status": stat})
    uses = [sum(1 for k in range(K) if k % NS == j) for j in range(NS)]
    conv = sum(per[j]["conv"] * uses[j] for j in range(NS)) / max(K, 1)          # converged instances per step, averaged over the timed steps
    status = np.concatenate([p["status"] for p in per])
    st = {"nfact": sum(p["nfact"] for p in per) / NS, "iters": sum(p["iters"] for p in per) / NS}
    launches_per_step = per[0]["launches"]                            # kernels (the two queue memsets are not kernels)
    kernel_ms = float(np.mean([p["kernel_ms"] for p in per]))
    extras = {}
    if not a.no_extras and a.config == 2:
        # ---------------------------------------------------------------- cold start of the same batch ("also report cold start")
        cms = []
        for _ in range(2):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            oc = solver.solve_device(*cur_t, mass_t, k1_t, 0, stream=stream.cuda_stream)
            c1.record(stream)
            torch.cuda.synchronize()
            cms.append(c0.elapsed_time(c1))
        cconv = int((oc["status"] == 0).sum().item())
        cst = solver.last_stats()
        extras["cold_start"] = {"value": cconv / (min(cms) * 1e-3), "unit": UNIT, "converged_fraction": cconv / B, "iters_per_solve": cst["iters"] / B,
                                "what": "the same batch from the solver's own initial guess (x_i = x0, f_z = m g / #contact vertices), this GPU"}
        # ---------------------------------------------------------------- rolling replay: consecutive ticks, each warm-started from the one before
        roll_t = [[t(x) for x in ticks[2 + r]] for r in range(R + 1)]
        solver.warm_restore(B, stream.cuda_stream)
        rconv, rit = 0, 0
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record(stream)
        outs = []
        for r in range(R):
            outs.append(solver.solve_device(*roll_t[r], mass_t, k1_t, WM, stream=stream.cuda_stream)["status"])
        r1.record(stream)
        torch.cuda.synchronize()
        rconv = int(sum((o == 0).sum().item() for o in outs))
        extras["rolling_replay"] = {"value": rconv / (r0.elapsed_time(r1) * 1e-3), "unit": UNIT, "ticks": R, "converged_fraction": rconv / (B * R),
                                    "what": "%d consecutive ticks t, t+1, ... of every instance, each warm-started (mode %d) from the device-resident solution of the "
                                            "tick before; the launch order follows the previous tick's work" % (R, WM)}
    # ------------------------------------------------------------------ single-instance latency (one robot, consecutive ticks of the walk)
    lat = None
    if rank == 0 and not a.no_extras and N in (10, 20):
        w = load_ticks(N)
        one = pkg.BatchSolver(N, 1, device=local, **over)
        ms = []
        T0, T1 = 100, 900                                                # standing, six steps, the push window's start
        for tk in range(T0, T1):
            args = [w[k][tk:tk + 1] for k in ("x0", "com_ref", "foot_ref", "gamma")]
            t0 = time.perf_counter()
            r1_ = one.solve_host(*args, float(w["mass"]), float(w["k1"]), 0 if tk == T0 else WM, traj_batch=1)   # what the drop-in class calls per tick
            if tk > T0:
                ms.append((time.perf_counter() - t0) * 1e3)
        one.close()
        lat = {"p50_ms": float(np.percentile(ms, 50)), "p90_ms": float(np.percentile(ms, 90)), "p99_ms": float(np.percentile(ms, 99)), "max_ms": float(np.max(ms)),
               "what": "cmpc_solve_host_traj of ONE instance (H2D + kernels + D2H of x1, u0 and the trajectories), ticks %d..%d of the recorded walk in order, "
                       "each warm-started from the tick before" % (T0 + 1, T1 - 1)}
    # ------------------------------------------------------------------ batched closed loop (Fleet: device assembly + solve + plant), robot-ticks/s
    fleet_line = None
    if rank == 0 and not a.no_extras and a.config == 2 and N in (10, 20):
        planner, com_ref, params, initial = walk_tables()
        params = dict(params, N=N)
        FB = min(B, 4096)
        rngf = np.random.default_rng(7)
        fleet = pkg.Fleet(FB, planner, params, com_ref, initial, hw_trace=initial["hw_meas"], device=local, tick_offset=rngf.integers(0, 1800, FB))
        wf = load_ticks(N)
        o = fleet.offset.cpu().numpy()
        fleet.com_pos, fleet.com_vel = t(wf["x0"][o, 0:3]), t(wf["x0"][o, 3:6])            # every robot starts from the recorded state of its own tick
        fleet.hw, fleet.theta = t(wf["x0"][o, 6:9]), t(wf["x0"][o, 9:12])
        for tk in range(3):
            fleet.step(tk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        FT = 10
        for tk in range(3, 3 + FT):
            fleet.step(tk)
        torch.cuda.synchronize()
        dtf = time.perf_counter() - t0
        fleet_line = {"value": FB * FT / dtf, "unit": "robot-ticks/s", "robots": FB, "ticks": FT, "alive_fraction": float(fleet.alive.double().mean().item()),
                      "what": "Fleet.step: cmpc_assemble_device + cmpc_solve_device (warm) + plant update + step-adjustment scatter, robots at random phases of the walk"}
        del fleet
    # ------------------------------------------------------------------ e2e through the host-buffer C-ABI call
    torch.cuda.set_stream(torch.cuda.default_stream(dev))

    def step_host(j):
        m = samples[j % NS]
        if warm_replay:
            m["solver"].warm_restore(B)
            return m["solver"].solve_host(*m["cur"], m["mass_h"], m["k1_h"], WM)
        return m["solver"].solve_host(*m["cur"], m["mass_h"], m["k1_h"], 0)

    for j in range(min(W, 2)):
        step_host(j)
    barrier()
    t0 = time.perf_counter()
    conv_e2e = 0
    for j in range(K):
        res = step_host(j)
        conv_e2e += int((res["status"] == 0).sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = B * (20 + 9 * N + 8 * N + 2 * (N + 1) + 2) * 8
    d2h = B * ((20 + 32 + 20 + 2) * 8 + 8)
    # ------------------------------------------------------------------ reduce over ranks
    if world > 1:
        v = torch.tensor([total_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        c = torch.tensor([conv, conv_e2e, st["nfact"], st["iters"]], device=dev, dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_ms, e2e_s = float(v[0]), float(v[1])
        conv_all, conv_e2e_all, nfact_all, iters_all = [float(x) for x in c]
    else:
        conv_all, conv_e2e_all, nfact_all, iters_all = conv, conv_e2e, st["nfact"], st["iters"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = conv_all * K / (total_ms * 1e-3)
    e2e_value = conv_e2e_all / e2e_s
    # ------------------------------------------------------------------ roofline of the solve kernel (rank 0's launch)
    fp64_peak = pkg.measure_fp64_peak(local)
    flops = st["nfact"] * N * F_FACT + st["iters"] * N * F_SOLVE
    tf = flops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = B * 8.0 * (123 * N + 62)
    gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of the solve kernel launches of ONE step, measured by ncu on this shape (scripts/dram_bytes.sh); null when
    # the committed measurement is for another shape
    traffic, traffic_note = None, "no ncu measurement committed for this shape (profiles/r02_dram_bytes.json)"
    try:
        m = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_bytes.json")))
        if m.get("horizon") == N and m.get("batch") == B and m.get("config") == a.config:
            traffic = float(m["dram_bytes_per_step"])
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum over the cmpc_solve_kernel launches of one step, ncu on this shape (%s): %.1fx the "
                            "algorithmic bytes" % (m.get("source", "profiles/r02_dram_bytes.json"), traffic / alg_bytes))
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "traffic": traffic,
                "kernel": "cmpc_solve_kernel", "kernel_ms": kernel_ms, "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "flops_convention": "dense: %d*N per Riccati factorisation x %d factorisations + %d*N per solve x %d iterations"
                                    % (F_FACT, st["nfact"], F_SOLVE, st["iters"]),
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "algorithmic_bytes_per_solve": 8 * (123 * N + 62), "traffic_note": traffic_note}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if a.config == 5 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (recorded surrogate walk, tests/golden/walk_ticks_N20.npz + walk_inputs.npz)", "config": config,
            "converged_fraction": conv_all / (B * world), "iters_per_solve": iters_all / (B * world),
            "factorisations_per_solve": nfact_all / (B * world), "status_histogram": (np.bincount(status, minlength=7) / NS).tolist(),
            "step_ms": [round(x, 3) for x in step_ms],
            "p50_batch_latency_ms": float(np.percentile(step_ms, 50)), "max_batch_latency_ms": float(np.max(step_ms)), "single_instance_latency": lat,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * K * world, "roofline": roofline, "clocks": clocks}
    line.update(extras)
    if fleet_line:
        line["fleet"] = fleet_line
    if world == 1 and not a.no_cpu_baseline:
        ncpu = os.cpu_count()
        nsamp = a.cpu_sample or 128 * ncpu
        Nr = N if N in (10, 20) else 20
        try:
            r = cpu_reference_rate(Nr, nsamp, seed=0)
            line["cpu_baseline"] = {"value": r["rate"], "unit": UNIT, "cores": r["procs"], "kind": "port",
                                    "sample": "%d instances of the N=%d tick replay (primal warm start from tick t-1 as the reference does, %.1f s wall, %d converged, %.1f "
                                              "iterations/solve, %.1f ms/solve/core), restated reference oracle/ipm_c.c at the reference's IPOPT settings (tol=1e-3, "
                                              "constr_viol_tol=1e-4; CasADi/IPOPT not installable offline), gcc -O3 -march=native, one process per core"
                                              % (nsamp, Nr, r["wall_s"], r["converged"], r["iters"], r["ms_per_solve_per_core"])}
            rt = cpu_reference_rate(Nr, max(nsamp // 2, ncpu), seed=0, tight=True)
            line["cpu_baseline_tight"] = {"value": rt["rate"], "unit": UNIT, "cores": rt["procs"], "kind": "port",
                                          "sample": "same oracle at the parity target's tolerance (1e-8 at mu=1e-9, oracle-T): %d instances, %.1f iterations/solve"
                                                    % (rt["n"], rt["iters"])}
        except Exception as e:  # the baseline is a reported extra: never lose the bench line over it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
