#!/bin/bash
mkdir -p gpurun_out
{
echo "== default (5 CTAs/SM) WCA"; timeout 200 python tools/bench_configs.py "C5 WCA"
for v in f4 f6; do echo "== $v WCA"; SADMC_GPU_LIB=$PWD/build_exp/$v.so timeout 200 python tools/bench_configs.py "C5 WCA"; done
} > gpurun_out/variants6.log 2>&1
timeout 200 python tools/sanitize_smoke.py 300 > gpurun_out/smoke_plain.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py 60 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 20 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
cat gpurun_out/variants6.log; cat gpurun_out/smoke_plain.log; tail -5 gpurun_out/sanitizer_memcheck.log; tail -5 gpurun_out/sanitizer_racecheck.log
