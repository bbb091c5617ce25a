cd "$GRAFT_REPO_ROOT"
timeout 400 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; tail -3 gpurun_out/bench_r02e.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r02e.json'))
print({k:j[k] for k in ('value','ms_per_step','iters_per_solve','converged_fraction','step_ms','gpu_launches')}, j['e2e']['value'], j['rolling_replay']['value'], j['fleet']['value'], j['roofline']['frac'], j['roofline']['kernel_ms'], j['cpu_baseline']['value'])
PY
timeout 200 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3:', j['value'], j['converged_fraction'])"
