#!/bin/bash
# throughput only (no parity) for experiment builds
for n in "$@"; do
  export SADMC_GPU_LIB=$PWD/build_exp/$n.so
  echo -n "$n: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4
done
