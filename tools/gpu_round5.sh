#!/bin/bash
mkdir -p gpurun_out
{
for v in g1u2 g1u6 g1u8; do echo "== $v lanes=1"; SADMC_GPU_LIB=$PWD/build_exp/$v.so timeout 200 python tools/profile_lj.py 75776 1 20000 4; done
echo "== default lanes=1"; timeout 200 python tools/profile_lj.py 75776 1 20000 4
for v in e2 e2u8; do echo "== $v lanes=2"; SADMC_GPU_LIB=$PWD/build_exp/$v.so timeout 200 python tools/profile_lj.py 75776 2 20000 4; done
for v in f3 f5; do echo "== $v WCA"; SADMC_GPU_LIB=$PWD/build_exp/$v.so timeout 200 python tools/bench_configs.py "C5 WCA"; done
} > gpurun_out/variants5.log 2>&1
cat gpurun_out/variants5.log
