#!/bin/bash
# round 2, GPU call 3: LJ parity with deferred bookkeeping + speed A/B, exact-DOS gate (range-stable walkers), WCA ncu capture
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_lj.py tests/test_gpu_resume.py -q) > gpurun_out/r02_pytest_call3.log 2>&1
tail -6 gpurun_out/r02_pytest_call3.log
for n in "" nodefer; do
  if [ -n "$n" ]; then export SADMC_GPU_LIB=$PWD/build_exp/$n.so; else unset SADMC_GPU_LIB; fi
  echo -n "lj31 ${n:-defer}: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4
  echo -n "lj31 ${n:-defer} (second run): "; timeout 300 python tools/profile_lj.py 75776 1 20000 4
done 2>&1 | tee gpurun_out/r02_lj_defer.log
unset SADMC_GPU_LIB
timeout 900 python tools/dos_gate.py --schedule 1e6,3e6,1e7 --systems linear,quadratic --out gpurun_out/r02_dos_gate_fake.jsonl > gpurun_out/r02_dos_gate_fake.log 2>&1
cut -c1-330 gpurun_out/r02_dos_gate_fake.log | tail -40
timeout 600 python tools/dos_gate.py --schedule 1e6,1e7 --systems two-wells --out gpurun_out/r02_dos_gate_two_wells.jsonl > gpurun_out/r02_dos_gate_two_wells.log 2>&1
cut -c1-330 gpurun_out/r02_dos_gate_two_wells.log | tail -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:move_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_wca_g8fast \
  python tools/bench_wca.py --variants g8fast --moves 2000 --burn-in 20000 --cpu-seconds 0 > gpurun_out/r02_wca_ncu.log 2>&1
tail -3 gpurun_out/r02_wca_ncu.log
