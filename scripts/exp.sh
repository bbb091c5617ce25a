cd "$GRAFT_REPO_ROOT"
for mi in 300 250; do
echo "### config 5 max_iter $mi"; timeout 900 python bench.py --config 5 --steps 2 --warmup 1 --no-cpu-baseline --cfg max_iter=$mi 2> gpurun_out/cfg5.err | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'converged_fraction', 'iters_per_solve', 'status_histogram')}, j['e2e']['value'])
"; tail -3 gpurun_out/cfg5.err
done
