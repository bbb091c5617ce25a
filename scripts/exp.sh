cd "$GRAFT_REPO_ROOT"
for c in 5 3; do
timeout 300 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/cfg$c.err | tee gpurun_out/bench_config${c}_1gpu_retry.json | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'converged_fraction', 'iters_per_solve', 'factorisations_per_solve', 'status_histogram')}, j['e2e']['value'])
"; done
