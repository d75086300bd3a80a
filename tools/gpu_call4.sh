#!/bin/bash
# round 2, GPU call 4: LJ31 variants round-robin, WCA after the branch-free rewrite, two-wells gate with residual dump
mkdir -p gpurun_out
tools/ab_lj.sh 3 main nodefer zigno nodefer_zigno 2>&1 | tee gpurun_out/r02_lj_ab1.log
(time timeout 900 python -m pytest tests/test_gpu_fluids.py -q -k wca) > gpurun_out/r02_pytest_call4.log 2>&1
tail -4 gpurun_out/r02_pytest_call4.log
timeout 600 python tools/bench_wca.py --variants g8fast,g16fast,g4fast,g8 --cpu-seconds 0 2>&1 | tee gpurun_out/r02_wca_variants2.jsonl
timeout 600 python tools/dos_gate.py --schedule 1e6,1e7 --systems two-wells --dump gpurun_out/r02_dos --out gpurun_out/r02_dos_gate_two_wells.jsonl > gpurun_out/r02_dos_gate_two_wells.log 2>&1
cut -c1-330 gpurun_out/r02_dos_gate_two_wells.log | tail -8
