#!/bin/bash
# Full GPU pass: parity tests, bench line, ncu launch list and one full capture of the solve kernel.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$TAG.json
if [ "$2" == "ncu" ]; then
  timeout 300 python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
  timeout 300 python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline > gpurun_out/plain2_$TAG.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
      python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
  ls -la gpurun_out/
fi
