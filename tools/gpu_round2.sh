#!/bin/bash
# Second GPU-box session: the whole parity suite (no -x), walker-count sweep of the headline kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
{
  for W in 18944 37888 56832 75776 113664 151552; do echo "W=$W lanes=1"; timeout 300 python tools/profile_lj.py $W 1 20000 4; done
} > gpurun_out/walkers.log 2>&1
tail -5 gpurun_out/pytest_gpu_full.log; cat gpurun_out/walkers.log
