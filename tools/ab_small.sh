#!/bin/bash
for n in ${VARIANTS:-main small_mb6 small_mb8 small_mb12}; do
  if [ "$n" = main ]; then unset SADMC_GPU_LIB; else export SADMC_GPU_LIB=$PWD/build_exp/$n.so; fi
  echo "== $n"; timeout 300 python tools/bench_configs.py "C1 ising32 SAD 262144" "C1 ising32 WL" "C2 fake linear" "C2 two-wells" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print('  %-55s %.4g moves/s' % (d['config'], d.get('moves_per_s', float('nan'))), d.get('error',''))
    except Exception: print(l.strip()[:200])
"
done
