#!/bin/bash
# round-2 GPU session C: A/B of the chain-shortening variants (same source, -D switches), then the parity suite
mkdir -p gpurun_out
for v in libvariant_base libvariant_nofma libvariant_fmaonly libsimplyp_b200; do
  echo "== $v"
  SIMPLYP_B200_LIB=$PWD/simplyp_b200/lib/$v.so python scripts/exp_minblocks.py 10000 40000 160000 2>&1 | grep -E "default"
done > gpurun_out/r2c_variants.txt 2>&1
cat gpurun_out/r2c_variants.txt
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
