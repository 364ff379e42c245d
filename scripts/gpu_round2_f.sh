#!/bin/bash
# round-2 GPU session F: member-chunk ordering of network launches — parity suite, then configs 3 and 5 with and without it
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
for cm in default 100000; do
  if [ $cm = default ]; then unset SIMPLYP_CHUNK_MEMBERS; else export SIMPLYP_CHUNK_MEMBERS=$cm; fi
  for spec in "3 64" "3 256" "5 4" "5 8"; do set -- $spec
    python bench.py --config $1 --members $2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk=$cm config $1 M=$2: %.1f ms  %.3e /s  fp64 %.3f  hbm %.1f GB/s'%(d['ms_per_step'],d['value'],d['roofline']['fp64']['frac'],d['roofline']['hbm']['achieved']))"
  done
done > gpurun_out/r2f_chunks.txt 2>&1
cat gpurun_out/r2f_chunks.txt
