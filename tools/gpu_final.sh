#!/bin/bash
# Round-end check on one B200: smoke, the whole parity suite, both bench arms, per-config throughput,
# racecheck of the kernels that were fixed, launch list of the bench, one ncu capture of the headline kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py 20 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:move_kernel -s 1 -c 1 -o gpurun_out/lj31_r01c python tools/profile_lj.py 75776 1 400 4 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu_full.log; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/sanitizer_racecheck.log
