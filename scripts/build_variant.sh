#!/bin/bash
# scripts/build_variant.sh NAME [nvcc flags...] -> scratch_libs/lib_NAME.so, prints ptxas resource usage of the solve kernel
cd "$(dirname "$0")/.." || exit 1
P=online-non-linear-centroidal-mpc-with-stability-guarantees-for-robust-locomotion-of-legged-robots-_b200
name=$1; shift
mkdir -p lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -diag-suppress 170 -Xptxas -v \
  -shared -Xcompiler -fPIC "$@" -o lib/variants/lib_$name.so $P/csrc/cmpc_kernels.cu 2>&1 | grep -A3 "Compiling entry function.*cmpc_solve_kernel" | grep -E "stack|Used" | tr '\n' ' ' | sed "s/^/$name: /"
echo
