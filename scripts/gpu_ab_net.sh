#!/bin/bash
# A/B of two builds on the network configurations (one session), then the GPU test suite on the default build
mkdir -p gpurun_out
VARS="$@"; for r in 1 2; do for v in $VARS; do
  for spec in "3 64" "5 4"; do set -- $spec
    SIMPLYP_B200_LIB=$PWD/simplyp_b200/lib/$v.so python bench.py --config $1 --members $2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v config $1 M=$2: %.1f ms  %.3e /s'%(d['ms_per_step'],d['value']))"
  done
done; done > gpurun_out/r2_ab_net.txt 2>&1
cat gpurun_out/r2_ab_net.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
