#!/bin/bash
# round 2, GPU call 5: deferred bookkeeping (single call site) vs not, WCA G=4 fix, DOS gate test, two-wells big step,
# then the LJ31 heat-capacity production run at the headline SAD parameters and the canonical cross-check
mkdir -p gpurun_out
tools/ab_lj.sh 2 main defer2 2>&1 | tee gpurun_out/r02_lj_ab2.log
(time timeout 900 python -m pytest tests/test_gpu_fluids.py tests/test_gpu_dos_gate.py tests/test_gpu_merge.py -q -k "wca or dos or fake or randomize") > gpurun_out/r02_pytest_call5.log 2>&1
tail -12 gpurun_out/r02_pytest_call5.log | cut -c1-250
timeout 300 python tools/bench_wca.py --variants g8fast,g4fast --cpu-seconds 0 2>&1 | tee gpurun_out/r02_wca_variants3.jsonl
timeout 600 python tools/dos_gate.py --schedule 1e6,1e7 --systems two-wells-bigstep --dump gpurun_out/r02_dos --out gpurun_out/r02_dos_gate_two_wells_bigstep.jsonl 2>&1 | cut -c1-330 | tail -6
timeout 1500 python tools/lj31_cv_run.py --walkers 37888 --schedule 1e7,3e7,1e8,2e8 --chunk 2e6 --settled 0 --tag lj31_cv_headline 2>&1 | tee gpurun_out/r02_lj31_cv_headline.log
timeout 600 python tools/lj31_canonical.py --temperatures 0.1,0.2,0.25,0.3,0.35,0.4 --walkers 18944 --moves 1e7 --out gpurun_out/r02_lj31_canonical.json 2>&1 | tee gpurun_out/r02_lj31_canonical.log
