#!/bin/bash
# DRAM traffic of the solve kernel on the BENCHED shape (batch 4096, N = 20, warm replay): ncu dram__bytes_read/write of
# every cmpc_solve_kernel launch of one timed step -> gpurun_out/dram_bytes_TAG.csv + profiles-ready JSON.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/plain_dram_$TAG.log 2>&1 || exit 1
# launches of the solve kernel in that command: 2 setup solves, then one launch per step (warm-up, timed, kernel-time repeat, e2e warm-up, e2e)
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cmpc_solve_kernel --csv \
    --log-file gpurun_out/dram_bytes_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_dram_$TAG.log 2>&1
python - "$TAG" <<'PY'
import csv, json, sys
tag = sys.argv[1]
rows = []
with open("gpurun_out/dram_bytes_%s.csv" % tag) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
by = {}
for r in rows:
    by.setdefault(int(r["ID"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
launch = []
for i in sorted(by):
    m = by[i]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    launch.append((rd, wr, m["gpu__time_duration.sum"][0]))
# setup: tick t-2 cold, tick t-1 warm; every launch after that is one step
steps = [launch[2 + k: 3 + k] for k in range(len(launch) - 2)]
per_step = [sum(a + b for a, b, _ in s) for s in steps]
out = {"config": 2, "batch": 4096, "horizon": 20, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, -k cmpc_solve_kernel, bench.py --steps 1 --warmup 1 (%s)" % tag,
       "launches": len(launch), "dram_bytes_per_step_all": per_step, "dram_bytes_per_step": sorted(per_step)[len(per_step) // 2] if per_step else None,
       "dram_read_write_first_step": [sum(a for a, _, _ in steps[0]), sum(b for _, b, _ in steps[0])] if steps else None}
json.dump(out, open("gpurun_out/r02_dram_bytes_%s.json" % tag, "w"), indent=1)
print(json.dumps(out))
PY
