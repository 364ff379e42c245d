#!/bin/bash
# round-2 session T: compute-sanitizer (memcheck, racecheck) on the epoch sweep and the three-resident plan
mkdir -p gpurun_out
python scripts/sanitizer_case.py > gpurun_out/r2t_plain.log 2>&1; tail -1 gpurun_out/r2t_plain.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitizer_case.py > gpurun_out/r2t_memcheck.log 2>&1; tail -3 gpurun_out/r2t_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitizer_case.py > gpurun_out/r2t_racecheck.log 2>&1; tail -3 gpurun_out/r2t_racecheck.log
