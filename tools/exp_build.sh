#!/bin/bash
# quick experiment build of the LJ31 thread-per-walker kernels only: tools/exp_build.sh NAME [extra nvcc flags]
set -e
name=$1; shift
cd "$(dirname "$0")/../sad_monte_carlo_b200"
mkdir -p ../build_exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall \
  -shared -cudart shared -DSADMC_EXPERIMENT_LJ31 -Xptxas -v "$@" -o ../build_exp/$name.so csrc/engine.cu 2>&1 | grep -A2 "move_kernelINS_11LjThreadSysILb1ELi31ELi1EEELi1E" | grep -v Compiling
