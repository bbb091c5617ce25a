#!/bin/bash
# Latency- or throughput-bound?  Same kernel, same 4096-instance batch, with 1, 2 and 3 resident CTAs per SM (shared-memory padding).
cd "$GRAFT_REPO_ROOT" || exit 1
for pad in 120000 40000 0; do
  echo "== smem pad $pad"
  CMPC_SMEM_PAD=$pad CMPC_LIB="$PWD/${1:-scratch_libs/lib_A.so}" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        j = json.loads(line)
        print({k: j[k] for k in ('value', 'ms_per_step', 'converged_fraction', 'iters_per_solve', 'factorisations_per_solve')}, j['roofline']['frac'])
    elif line: print(line[-300:])
"
done
