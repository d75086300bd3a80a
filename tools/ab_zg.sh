#!/bin/bash
# A/B of the two LJ31 move-kernel layouts in the in-tree library: z streamed from L2 (3 CTAs per SM; flags 4|128) against all
# coordinates in shared memory (2 CTAs per SM; flags 4|64), whole waves for both (113 664 walkers = 2 x 148 x 384 = 3 x 148 x 256),
# round-robin; then bench-like launches (200 000 moves after 1e6 burn-in) at one wave of each.
#   tools/profile_lj.py WALKERS LANES MOVES FLAGS REPS BURN_IN
for r in 1 2; do
  echo -n "round $r stream: "; timeout 300 python tools/profile_lj.py 113664 1 20000 132 4 200000 2>&1 | tail -1
  echo -n "round $r smem:   "; timeout 300 python tools/profile_lj.py 113664 1 20000 68 4 200000 2>&1 | tail -1
done
echo -n "stream 56832 long: "; timeout 300 python tools/profile_lj.py 56832 1 200000 132 3 1000000 2>&1 | tail -1
echo -n "smem 75776 long:   "; timeout 300 python tools/profile_lj.py 75776 1 200000 68 3 1000000 2>&1 | tail -1
