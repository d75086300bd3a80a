#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fluids.py tests/test_gpu_checkpoint.py tests/test_gpu_resume.py tests/test_golden.py -m gpu -q --durations=8 ) > gpurun_out/pytest_fluids.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fluids.log
timeout 600 python tools/bench_configs.py C5 > gpurun_out/configs_c5.jsonl 2>&1
tail -4 gpurun_out/pytest_fluids.log; cat gpurun_out/configs_c5.jsonl
