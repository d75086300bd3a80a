#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_host_cpp.py -m gpu -q --durations=5 ) > gpurun_out/pytest_host_cpp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_host_cpp.log
tail -40 gpurun_out/pytest_host_cpp.log
