#!/bin/bash
# Quick experiment build: recompile only the LJ thread-per-walker tolerance-tier unit with extra flags and link it
# with the cached objects of the regular build (python -m sad_monte_carlo_b200.build first).
#   [UNIT=kernels_lj_thread_fast_multi] tools/exp_build.sh NAME [extra nvcc flags]      ->  build_exp/NAME.so   (use with SADMC_GPU_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../sad_monte_carlo_b200"
mkdir -p ../build_exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall \
  -Xptxas -v "$@" -c -o ../build_exp/$name.o csrc/${UNIT:-kernels_lj_thread_fast}.cu 2>&1 | grep -A3 "11move_kernelINS_11LjThreadSysILb1ELi31ELi[124]ELi0ELb[01]EEELi1E\|LjPairedSysILi31EEELi1E" | grep -v Compiling || true
objs=$(ls csrc/_obj/*.o | grep -v "/${UNIT:-kernels_lj_thread_fast}.o")
nvcc -shared -cudart shared -gencode arch=compute_100a,code=sm_100a -o ../build_exp/$name.so $objs ../build_exp/$name.o
