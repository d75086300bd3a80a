#!/bin/bash
# Occupancy probe for the LJ thread-per-walker kernel: smaller clusters (LJ20, LJ16) leave shared memory for 3 / 4 CTAs per SM,
# so the same kernel can be timed at 8, 12 and 16 warps per SM (249 / 168 / 128 registers); LJ31 with three 96-thread CTAs (9 warps).
# Walker counts are whole waves for every variant compared.  The libraries (after python -m sad_monte_carlo_b200.build):
#   tools/exp_build.sh n20_2 -DSADMC_EXP_NT=20
#   tools/exp_build.sh n20_3 -DSADMC_EXP_NT=20 -DSADMC_LJT_BLOCK=128 -DSADMC_LJT_MIN_BLOCKS=3
#   tools/exp_build.sh n16_2 -DSADMC_EXP_NT=16
#   tools/exp_build.sh n16_4 -DSADMC_EXP_NT=16 -DSADMC_LJT_BLOCK=128 -DSADMC_LJT_MIN_BLOCKS=4
#   tools/exp_build.sh lj31_96x3 -DSADMC_LJT_BLOCK=96 -DSADMC_LJT_MIN_BLOCKS=3      (faults: the 96-thread shape is not supported)
# (each is a whole library of ~60 MB; gpurun refuses snapshots above 512 MiB, so build what one call needs)
run() { lib=$1; shift; echo -n "$lib: "; SADMC_GPU_LIB=$PWD/build_exp/$lib.so timeout 300 python "$@" 2>&1 | tail -1; }
run n20_2 tools/profile_ljn.py 20 -77.3 113664
run n20_3 tools/profile_ljn.py 20 -77.3 113664
run n16_2 tools/profile_ljn.py 16 -56.9 113664
run n16_2 tools/profile_ljn.py 16 -56.9 75776
run n16_4 tools/profile_ljn.py 16 -56.9 75776
run lj31_96x3 tools/profile_lj.py 85248 1 20000 4 4 200000
echo -n "main: "; timeout 300 python tools/profile_lj.py 75776 1 20000 4 4 200000
