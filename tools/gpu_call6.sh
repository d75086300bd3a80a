#!/bin/bash
# round 2, GPU call 6: fixed-weight production runs for the LJ31 heat capacity (two multicanonical iterations), WCA walker sweep + ncu
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_merge.py -q -k fixed_weights) 2>&1 | tail -3
timeout 900 python tools/lj31_production.py --weights tests/golden/lj31_cv_run_r02/lj31_cv_headline_2e+08.npz --moves 4e7 --out gpurun_out/lj31_production_1.npz 2>&1 | tail -2
timeout 900 python tools/lj31_production.py --weights gpurun_out/lj31_production_1.npz --moves 4e7 --out gpurun_out/lj31_production_2.npz 2>&1 | tail -2
for w in 9472 14208 18944; do timeout 300 python tools/bench_wca.py --variants g8fast --walkers $w --cpu-seconds 0; done 2>&1 | tee gpurun_out/r02_wca_walkers.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:move_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_wca_g8fast_v2 \
  python tools/bench_wca.py --variants g8fast --moves 2000 --burn-in 20000 --cpu-seconds 0 > gpurun_out/r02_wca_ncu2.log 2>&1
tail -2 gpurun_out/r02_wca_ncu2.log
