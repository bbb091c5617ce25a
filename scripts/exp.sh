cd "$GRAFT_REPO_ROOT"
# the other BASELINE configurations on one GPU + the phase counters of the final source (-DCMPC_PROFILE build)
for c in 3 4 5; do
timeout 300 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/cfg$c.err | tee gpurun_out/bench_config${c}_1gpu.json | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: j[k] for k in ('value', 'ms_per_step', 'converged_fraction', 'iters_per_solve', 'factorisations_per_solve', 'status_histogram')}, j['e2e']['value'])
"; done
python scripts/prof_ab.py lib/variants/lib_prof_final.so | tee gpurun_out/prof_final.json
CMPC_LIB=$PWD/lib/variants/lib_prof_final.so timeout 120 python scripts/phase_profile.py 20 4096 | tail -1 > gpurun_out/phase_final.json; cat gpurun_out/phase_final.json
timeout 200 python bench.py --horizon 10 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_N10.json | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=10', j['value'], j['e2e']['value'], j['single_instance_latency'])
"
