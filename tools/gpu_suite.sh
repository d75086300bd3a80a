#!/bin/bash
# smoke + the whole -m gpu suite on one B200 (what the driver runs at round end)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -2 gpurun_out/smoke.log; tail -16 gpurun_out/pytest_gpu_full.log
