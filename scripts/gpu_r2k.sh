#!/bin/bash
# round-2 session K: placement plan with 3 resident blocks per SM — A/B, timeline, invariance tests
mkdir -p gpurun_out
python scripts/exp_plan3.py 10000 11000 12000 13000 14000 > gpurun_out/r2k_plan3.log 2>&1; cat gpurun_out/r2k_plan3.log
SIMPLYP_B200_LIB=build/exp/libsimplyp_timeline.so python scripts/exp_timeline.py 10000 > gpurun_out/r2k_timeline.log 2>&1; tail -2 gpurun_out/r2k_timeline.log
python -m pytest tests -m gpu -x -q -k "planned or bench_ensemble or full_size_ensemble" 2>&1 | tail -5 > gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest.log
