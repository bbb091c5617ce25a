#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 900 python scripts/first_timing.py 2>&1 | tee gpurun_out/first_timing.log | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(golden_parity_cold and 10) or k2 or warm or device_entry or order or infeasible or capacity" 2>&1 | tail -15
