#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lj.py tests/test_golden.py tests/test_gpu_resume.py -m gpu -q -x --durations=5 ) > gpurun_out/pytest_lj.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lj.log
{
echo "== predraw lanes=1"; timeout 200 python tools/profile_lj.py 75776 1 20000 4
echo "== no predraw lanes=1"; SADMC_GPU_LIB=$PWD/build_exp/nopre.so timeout 200 python tools/profile_lj.py 75776 1 20000 4
echo "== predraw lanes=1 again"; timeout 200 python tools/profile_lj.py 75776 1 20000 4
} > gpurun_out/variants7.log 2>&1
tail -6 gpurun_out/pytest_lj.log; cat gpurun_out/variants7.log
