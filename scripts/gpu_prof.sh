#!/bin/bash
# phase profile (CMPC_PROFILE build) + one ncu --set full capture of the solve kernel (batch 444 = 3 CTAs x 148 SMs)
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r01e}
CMPC_LIB=$PWD/scratch_libs/lib_pk3prof.so python scripts/phase_profile.py 20 4096 2>&1 | tail -1 | tee gpurun_out/phase_$TAG.json
timeout 300 python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/ | tail -5
