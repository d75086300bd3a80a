#!/bin/bash
# round 2, GPU call 2: WCA lane-group kernels (parity + throughput), merge tests, exact-DOS gate with the settled filter
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_fluids.py tests/test_gpu_merge.py -q -k "wca or merge or randomize or verify or fold or shard") > gpurun_out/r02_pytest_call2.log 2>&1
tail -15 gpurun_out/r02_pytest_call2.log
timeout 600 python tools/bench_wca.py --cpu-seconds 8 2>&1 | tee gpurun_out/r02_wca_variants.jsonl
timeout 300 python tools/bench_wca.py --variants g8fast,g4fast --walkers 18944 --cpu-seconds 0 2>&1 | tee -a gpurun_out/r02_wca_variants.jsonl
timeout 900 python tools/dos_gate.py --schedule 1e6,3e6 --systems linear,quadratic --dump gpurun_out/r02_dos --out gpurun_out/r02_dos_gate_explore2.jsonl > gpurun_out/r02_dos_gate_explore2.log 2>&1
cut -c1-400 gpurun_out/r02_dos_gate_explore2.log | tail -30
