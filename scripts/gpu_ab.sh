#!/bin/bash
# A/B of two builds of the library in ONE session (same box): alternate them, three rounds
mkdir -p gpurun_out
for r in 1 2 3; do for v in "$@"; do
  echo "== $v (round $r)"; SIMPLYP_B200_LIB=$PWD/simplyp_b200/lib/$v.so python scripts/exp_minblocks.py 10000 160000 2>&1 | grep default
done; done > gpurun_out/r2_ab.txt 2>&1
cat gpurun_out/r2_ab.txt
