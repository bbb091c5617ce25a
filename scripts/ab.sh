#!/bin/bash
# A/B timing of several builds of the library on ONE box: scripts/ab.sh libA.so libB.so ...  (paths relative to the repo root)
# AB_ARGS: extra bench.py arguments; a "name=path" argument labels the line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for so in "$@"; do
  echo "== $so $AB_ARGS"
  CMPC_LIB="$PWD/$so" timeout 600 python bench.py --steps ${AB_STEPS:-8} --warmup 3 --no-cpu-baseline --no-extras $AB_ARGS 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        j = json.loads(line)
        print({k: j[k] for k in ('value', 'ms_per_step', 'converged_fraction', 'iters_per_solve', 'factorisations_per_solve')}, j['roofline']['frac'])
    elif line: print(line[-300:])
"
done
