#!/bin/bash
# Full GPU pass: parity tests, bench line, optionally the ncu launch list + one full capture of the solve kernel + DRAM bytes.
#   scripts/gpu_suite.sh TAG [ncu]
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "$2" == "ncu" ]; then
  bash scripts/dram_bytes.sh $TAG
  timeout 300 python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline --no-extras > gpurun_out/plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$TAG.log 2>&1
  timeout 300 python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline --no-extras > gpurun_out/plain2_$TAG.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cmpc_solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
      python bench.py --steps 1 --warmup 1 --batch 444 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1
  ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>&1
  ls -la gpurun_out/
fi
