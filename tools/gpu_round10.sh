#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 900 python -m pytest tests/test_gpu_lj.py -m gpu -q --durations=5 ) > gpurun_out/pytest_lj.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lj.log
{
echo "== lanes=2 W=85248"; timeout 200 python tools/profile_lj.py 85248 2 20000 4
echo "== lanes=1 W=75776"; timeout 200 python tools/profile_lj.py 75776 1 20000 4
echo "== lanes=4 W=75776"; timeout 200 python tools/profile_lj.py 75776 4 20000 4
} > gpurun_out/variants10.log 2>&1
tail -3 gpurun_out/smoke.log; tail -8 gpurun_out/pytest_lj.log; cat gpurun_out/variants10.log
