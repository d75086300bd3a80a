#!/bin/bash
mkdir -p gpurun_out
{
echo "== default lanes=1"; timeout 200 python tools/profile_lj.py 75776 1 20000 4
for v in e2u4 e2u8 e2u16; do echo "== $v lanes=2"; SADMC_GPU_LIB=$PWD/build_exp/$v.so timeout 200 python tools/profile_lj.py 75776 2 20000 4; done
echo "== e2u8 lanes=2 W=113664"; SADMC_GPU_LIB=$PWD/build_exp/e2u8.so timeout 200 python tools/profile_lj.py 113664 2 20000 4
} > gpurun_out/variants8.log 2>&1
timeout 600 python tools/bench_configs.py C1 C2 > gpurun_out/configs_c12.jsonl 2>&1
cat gpurun_out/variants8.log gpurun_out/configs_c12.jsonl
