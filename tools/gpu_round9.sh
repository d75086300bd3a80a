#!/bin/bash
mkdir -p gpurun_out
{
for W in 56832 85248 113664; do echo "== e2u16 lanes=2 W=$W"; SADMC_GPU_LIB=$PWD/build_exp/e2u16.so timeout 200 python tools/profile_lj.py $W 2 20000 4; done
echo "== default lanes=1 W=113664"; timeout 200 python tools/profile_lj.py 113664 1 20000 4
} > gpurun_out/variants9.log 2>&1
cat gpurun_out/variants9.log
