#!/bin/bash
# on the GPU box: parity subset + throughput for each experiment build: tools/exp_run.sh a b c
for n in "$@"; do
  export SADMC_GPU_LIB=$PWD/build_exp/$n.so
  echo "=== $n"
  timeout 600 python -m pytest tests/test_gpu_lj.py -x -q -k "thread_per_walker and (31 or fast)" 2>&1 | tail -3
  timeout 300 python tools/profile_lj.py 75776 1 20000 4
done
