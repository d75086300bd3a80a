#!/bin/bash
# One GPU-box session: parity suite, bench (both arms), per-config throughput, lane variants, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
{
  echo "lanes=1 default"; timeout 300 python tools/profile_lj.py 75776 1 20000 4
  echo "lanes=2 default (4 CTAs/SM, 128 regs)"; timeout 300 python tools/profile_lj.py 75776 2 20000 4
  echo "lanes=4 default"; timeout 300 python tools/profile_lj.py 75776 4 20000 4
  export SADMC_GPU_LIB=$PWD/build_exp/e2.so
  echo "lanes=2 e2 (3 CTAs/SM, 168 regs)"; timeout 300 python tools/profile_lj.py 75776 2 20000 4
  echo "lanes=4 e2"; timeout 300 python tools/profile_lj.py 75776 4 20000 4
  unset SADMC_GPU_LIB
} > gpurun_out/lanes.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:move_kernel -s 1 -c 1 -o gpurun_out/lj31_r01b python tools/profile_lj.py 75776 1 400 4 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json; cat gpurun_out/lanes.log
