#!/bin/bash
# A/B timing of LJ31 kernel variants, round-robin so that clock / thermal drift hits all variants alike:
#   tools/ab_lj.sh ROUNDS name1 name2 ...    ("main" = the in-tree library, others = build_exp/NAME.so)
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for n in "$@"; do
    if [ "$n" = main ]; then unset SADMC_GPU_LIB; else export SADMC_GPU_LIB=$PWD/build_exp/$n.so; fi
    echo -n "round $r $n: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4 6 200000
  done
done
