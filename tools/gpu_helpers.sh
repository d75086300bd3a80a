#!/bin/bash
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_lj_helpers.py -m gpu -q -x ) > gpurun_out/pytest_helpers.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_helpers.log
( echo "== helper warps (flags 12) W=75776"; timeout 40 python tools/profile_lj.py 75776 1 20000 12 ) > gpurun_out/helpers_speed.log 2>&1
tail -15 gpurun_out/pytest_helpers.log; cat gpurun_out/helpers_speed.log
