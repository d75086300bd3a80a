#!/bin/bash
# A/B: LJ31 with z streamed from L2 (in-tree, 3 CTAs per SM) against the all-shared-memory layout (build_exp/smemz.so, 2 CTAs per SM),
# whole waves for both (113 664 walkers = 2 x 148 x 384 = 3 x 148 x 256), round-robin.
for r in 1 2; do
  echo -n "round $r zg: "; timeout 300 python tools/profile_lj.py 113664 1 20000 4 4 200000 2>&1 | tail -1
  echo -n "round $r smemz: "; SADMC_GPU_LIB=$PWD/build_exp/smemz.so timeout 300 python tools/profile_lj.py 113664 1 20000 4 4 200000 2>&1 | tail -1
done
echo -n "zg 75776: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4 4 200000 2>&1 | tail -1
echo -n "zg 56832: "; timeout 300 python tools/profile_lj.py 56832 1 20000 4 4 200000 2>&1 | tail -1
