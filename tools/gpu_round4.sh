#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:move_kernel -s 1 -c 1 -o gpurun_out/wca_r01 python tools/bench_configs.py "C5 WCA" > gpurun_out/ncu_wca.log 2>&1
tail -3 gpurun_out/ncu_wca.log
