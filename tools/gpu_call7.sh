#!/bin/bash
# round 2, GPU call 7: full GPU test suite, WCA after the extras prefetch, C4/C5 script smoke on one GPU, bench.py, LJ31 ncu capture
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log | cut -c1-300
timeout 300 python tools/bench_wca.py --variants g8fast,g16fast --walkers 18944 --cpu-seconds 8 2>&1 | tee gpurun_out/r02_wca_final.jsonl
timeout 600 python tools/run_c4_c5.py --lj-walkers 4736 --lj-moves 2e5 --wca-walkers 2368 --wca-moves 2e4 --out gpurun_out/r02_c4_c5_smoke 2>&1 | cut -c1-600 | tail -8
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-1500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:move_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_lj31_main \
  python tools/profile_lj.py 75776 1 20000 4 3 1000000 > gpurun_out/r02_lj31_ncu.log 2>&1
tail -2 gpurun_out/r02_lj31_ncu.log
python tools/measure_fp64_peak.py 2>&1 | tail -1
