#!/bin/bash
# round 2, GPU call 1: new merge/halting tests, exact-DOS gate exploration, LJ31 ablations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
(time timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_ising.py tests/test_gpu_engine_extras.py tests/test_gpu_fake.py -x -q) > gpurun_out/r02_pytest_call1.log 2>&1
tail -5 gpurun_out/r02_pytest_call1.log
timeout 900 python tools/dos_gate.py --schedule 1e5,1e6,3e6 --out gpurun_out/r02_dos_gate_explore.jsonl > gpurun_out/r02_dos_gate_explore.log 2>&1
tail -30 gpurun_out/r02_dos_gate_explore.log
for n in base blk288 blk256 noload nort norecomp noslow; do
  export SADMC_GPU_LIB=$PWD/build_exp/$n.so
  echo -n "$n: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4
done 2>&1 | tee gpurun_out/r02_lj_ablations.log
