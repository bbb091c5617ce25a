#!/bin/bash
# Multi-GPU bench lines (one process per GPU, torchrun): scripts/multi_gpu.sh NGPUS "2 3 5"
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
G=${1:-2}; CFGS=${2:-"2 3 5"}
for c in $CFGS; do
  echo "### config $c on $G GPUs"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --config $c --steps 3 --warmup 2 \
      > gpurun_out/bench_config${c}_${G}gpu.json 2> gpurun_out/bench_config${c}_${G}gpu.err
  echo "rc=$?"; tail -2 gpurun_out/bench_config${c}_${G}gpu.err
  python - "gpurun_out/bench_config${c}_${G}gpu.json" <<'PY'
import sys, json
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith("{"):
        j = json.loads(line)
        print({k: j.get(k) for k in ("value", "n_gpus", "ms_per_step", "converged_fraction", "iters_per_solve", "scaling")}, "e2e", j["e2e"]["value"], j["config"]["batch_per_gpu"])
PY
done
