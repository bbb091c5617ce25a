#!/usr/bin/env python
"""bench.py -- converged centroidal-MPC solves/s on recorded-walk batches (BASELINE.json configs[1]).

A "step" is one MPC tick for a batch of independent instances: batch 4096 per GPU, horizon N = 20, instances =
ticks of the recorded surrogate walk sampled with replacement (seed = rank), each warm-started from the
solution of ITS previous tick, which is resident in the solver handle on the device (the solver's normal
operating mode: states, inputs, costates, slacks and multipliers stay in HBM across ticks).  Every timed step
first restores that previous-tick state from a device snapshot (device-to-device copy, inside the timed
region) and then solves the tick.

  value   converged solves/s with inputs resident in HBM (cmpc_solve_device), CUDA events, max over ranks
  e2e     same metric through the host-buffer C-ABI call the drop-in class uses (cmpc_solve_host): pinned
          staging + H2D + kernel + D2H inside the timed region
  roofline  FP64: dense-convention flops (121,749 N per Riccati factorisation + 10,368 N per
          solve) / kernel time vs the DFMA peak measured live on this GPU; HBM: algorithmic bytes
          8 (123 N + 62) per solve vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the restated reference (oracle/: CasADi/IPOPT cannot be installed here) on the host cores,
          one process per core, bounded sample of the same workload

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


def load_workload(N, batch, seed):
    w = np.load(os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N))
    rng = np.random.default_rng(seed)
    idx = rng.integers(2, len(w["x0"]), batch)
    take = lambda ii: (np.ascontiguousarray(w["x0"][ii]), np.ascontiguousarray(w["com_ref"][ii]),
                       np.ascontiguousarray(w["foot_ref"][ii]), np.ascontiguousarray(w["gamma"][ii]))
    return take(idx - 2), take(idx - 1), take(idx), float(w["mass"]), float(w["k1"]), idx


# ----------------------------------------------------------------------------------------------- CPU baseline
def _oracle_one(args):
    """One tick solved the way the reference does it: tick t-1 first (untimed), then tick t warm-started from the
    previous tick's primal solution (`opt.set_initial`, MPC file :630-631; slacks / multipliers / barrier restart)."""
    N, prev, cur, mass, k1 = args
    try:
        from oracle import ipm_c as orc
    except ImportError:
        from oracle import ipm_py as orc
        t0 = time.perf_counter()
        r = orc.solve_packed(N, *cur, mass, k1)
        return {"status": r["status"], "iters": r["iters"], "secs": time.perf_counter() - t0}
    r0 = orc.solve_packed(N, *prev, mass, k1)
    warm = (r0["X"], r0["U"]) if r0["status"] == 0 else None
    t0 = time.perf_counter()
    r = orc.solve_packed(N, *cur, mass, k1, warm=warm)
    if r["status"] != 0 and warm is not None:
        r = orc.solve_packed(N, *cur, mass, k1)
    return {"status": r["status"], "iters": r["iters"], "secs": time.perf_counter() - t0}


def oracle_name():
    try:
        from oracle import ipm_c  # noqa: F401
        return "oracle/ipm_c (C)"
    except ImportError:
        return "oracle/ipm_py (numpy/scipy, dense LDL')"


def cpu_reference_rate(N, n_instances, seed=0, procs=None):
    """Restated reference (oracle) on the host cores, one single-threaded process per core."""
    from multiprocessing import get_context
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    try:
        from oracle import ipm_c
        ipm_c.build()
    except ImportError:
        pass
    procs = procs or os.cpu_count()
    _, prev, cur, mass, k1, idx = load_workload(N, n_instances, seed)
    jobs = [(N, tuple(a[b] for a in prev), tuple(a[b] for a in cur), mass, k1) for b in range(n_instances)]
    with get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_oracle_one, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    conv = sum(1 for r in res if r["status"] == 0)
    busy = sum(r["secs"] for r in res)                 # CPU-seconds spent in the timed (tick t) solves, all cores busy concurrently
    dt = busy / procs
    return conv / dt, dt, conv, procs, float(np.mean([r["iters"] for r in res]))


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
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU baseline sample (0 = 128 per core, about 10-15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg", action="append", default=[], help="solver option override key=value (experiments only; the default run uses the library defaults)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    N, B, K, W = a.horizon, a.batch, a.steps, max(a.warmup, 0)
    config = {"workload": "replay of recorded surrogate-walk ticks, batch %d per GPU, horizon N=%d, full warm start from the "
                          "previous tick's device-resident solution" % (B, N),
              "batch_per_gpu": B, "horizon": N, "warm_start": "full (device snapshot of tick t-1, restored every step)",
              "launches_per_step": "cmpc_order_kernel (launch order from tick t-1's work) + cmpc_solve_kernel",
              "cache": "workspace %.1f GB per GPU streamed every iteration (>> 126 MB L2); no extra flush",
              "parallelism": "instances sharded over %d GPU(s), no collective on the hot path" % world}

    # ------------------------------------------------------------------ reference arm: restated CPU path only
    if a.impl == "reference":
        if rank != 0:
            return
        ncpu = os.cpu_count()
        per_step = a.cpu_sample or 32 * ncpu
        rates = []
        for s in range(W + K):
            r, dt, conv, procs, its = cpu_reference_rate(N, per_step, seed=s)
            if s >= W:
                rates.append((r, dt, conv))
        value = sum(c for _, _, c in rates) / sum(d for _, d, _ in rates)
        config["cache"] = "n/a (CPU)"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * np.mean([d for _, d, _ in rates]), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (recorded surrogate walk)", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                                 "sample": "%d instances per step of the same N=%d tick replay, primal warm start from tick t-1, restated "
                                           "reference (oracle/ipm_c: CasADi/IPOPT not installable offline), one process per core" % (per_step, N)},
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
    prev2, prev, cur, mass, k1, idx = load_workload(N, B, seed=rank)
    over = {k: float(v) for k, v in (kv.split("=") for kv in a.cfg)}
    solver = pkg.BatchSolver(N, B, device=local, **over)
    fp = solver.footprint()
    config["cache"] = config["cache"] % (B * fp["work_bytes_per_instance"] / 1e9)
    t = lambda x: torch.as_tensor(x, device=dev)
    mass_t, k1_t = t(np.full(B, mass)), t(np.full(B, k1))
    prev2_t, prev_t, cur_t = [t(x) for x in prev2], [t(x) for x in prev], [t(x) for x in cur]
    # the loop's steady state, untimed: tick t-2 from cold, tick t-1 warm-started from it.  The snapshot then holds what a
    # running controller has on the device when tick t arrives: the iterate of t-1 and the work its (warm) solve took,
    # which orders the launch of tick t (longest expected first)
    solver.solve_device(*prev2_t, mass_t, k1_t, 0)
    out = solver.solve_device(*prev_t, mass_t, k1_t, 2)
    torch.cuda.synchronize()
    solver.warm_save(B)
    stream = torch.cuda.Stream(dev)            # the solver's launches and the timing events share this stream
    torch.cuda.set_stream(stream)

    def step_device():
        solver.warm_restore(B, stream.cuda_stream)
        solver.solve_device(*cur_t, mass_t, k1_t, 2, out=out, stream=stream.cuda_stream)

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
    kernel_ms, conv_total, nfact, iters = [], 0, 0, 0
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]      # per-step marks inside the one timed region
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
    status = out["status"].cpu().numpy()
    conv = int((status == 0).sum())
    st = solver.last_stats()                                            # last step's kernel (all steps are identical work)
    # per-kernel duration of a step, measured live with CUDA events on the launching stream
    kms = []
    for _ in range(min(K, 3)):
        solver.warm_restore(B, stream.cuda_stream)
        solver.solve_device(*cur_t, mass_t, k1_t, 2, out=out, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        kms.append(solver.last_stats()["kernel_ms"])
    kernel_ms = float(np.mean(kms))
    # ------------------------------------------------------------------ single-instance latency (one robot, one tick)
    lat = None
    if rank == 0:
        one = pkg.BatchSolver(N, 1, device=local, **over)
        ms = []
        for k in range(48):                                              # 48 different ticks, each warm-started from its previous tick
            sl = slice(k, k + 1)
            one.solve_host(*[x[sl] for x in prev], mass, k1, 0)
            t0 = time.perf_counter()
            r1 = one.solve_host(*[x[sl] for x in cur], mass, k1, 2)      # host buffers in, host buffers out: what the drop-in class does
            ms.append((time.perf_counter() - t0) * 1e3)
        one.close()
        lat = {"p50_ms": float(np.percentile(ms, 50)), "p90_ms": float(np.percentile(ms, 90)), "max_ms": float(np.max(ms)),
               "what": "cmpc_solve_host of ONE instance (H2D + kernel + D2H), full warm start, 48 ticks of the replay"}
    # ------------------------------------------------------------------ e2e through the host-buffer C-ABI call
    for _ in range(min(W, 2)):
        solver.warm_restore(B); res = solver.solve_host(*cur, mass, k1, 2)
    barrier()
    t0 = time.perf_counter()
    conv_e2e = 0
    for _ in range(K):
        solver.warm_restore(B)
        res = solver.solve_host(*cur, mass, k1, 2)
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
    # DRAM bytes of the solve kernel per instance from the committed ncu capture (profiles/r01_ncu_summary.md, r01i:
    # 1.41 GB read + 2.05 GB written by a launch of 444 instances at N = 20), scaled to this launch; None for other horizons
    traffic = 7.79e6 * B if N == 20 else None
    roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "traffic": traffic,
                "kernel": "cmpc_solve_kernel", "kernel_ms": kernel_ms, "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "flops_convention": "dense: %d*N per Riccati factorisation x %d factorisations + %d*N per solve x %d iterations"
                                    % (F_FACT, st["nfact"], F_SOLVE, st["iters"]),
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "algorithmic_bytes_per_solve": 8 * (123 * N + 62),
                        "traffic_note": "roofline.traffic = DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the "
                                        "r01i capture, per instance, x batch): ~390x the algorithmic bytes by design -- per-stage records "
                                        "and factors are streamed through L2 / HBM every iteration -- and 1.4 % of the HBM bandwidth"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (recorded surrogate walk, tests/golden/walk_ticks_N%d.npz)" % N, "config": config,
            "converged_fraction": conv_all / (B * world), "iters_per_solve": iters_all / (B * world),
            "factorisations_per_solve": nfact_all / (B * world),
            "p50_batch_latency_ms": float(np.percentile(step_ms, 50)), "max_batch_latency_ms": float(np.max(step_ms)), "single_instance_latency": lat,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 2 * K * world, "roofline": roofline, "clocks": clocks}
    if world == 1 and not a.no_cpu_baseline:
        ncpu = os.cpu_count()
        nsamp = a.cpu_sample or 128 * ncpu
        try:
            r, dt, cconv, procs, its = cpu_reference_rate(N, nsamp, seed=0)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": procs, "kind": "port",
                                    "sample": "%d instances of the same N=%d tick replay (primal warm start from tick t-1 as the reference does, %.1f s per core, %d converged, %.1f iterations/solve), "
                                              "restated reference %s (CasADi/IPOPT not installable offline), one process per core"
                                              % (nsamp, N, dt, cconv, its, oracle_name())}
        except Exception as e:  # the baseline is a reported extra: never lose the bench line over it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
