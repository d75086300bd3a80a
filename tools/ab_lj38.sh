#!/bin/bash
# A/B of LJ38 kernel variants (CTA shape): tools/ab_lj38.sh main NAME ...   ("main" = the in-tree library, others = build_exp/NAME.so)
for n in "$@"; do
  if [ "$n" = main ]; then unset SADMC_GPU_LIB; else export SADMC_GPU_LIB=$PWD/build_exp/$n.so; fi
  echo "== $n"; timeout 300 python tools/bench_configs.py "C4 LJ38" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print('  %-45s %.4g moves/s' % (d['config'], d.get('moves_per_s', float('nan'))), d.get('error',''))
    except Exception: print(l.strip()[:200])
"
done
