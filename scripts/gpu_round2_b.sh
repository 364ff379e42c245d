#!/bin/bash
# round-2 GPU session B: parity suite on the 581-instruction build, bench config 2, ncu launch list + full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; head -c 300 gpurun_out/r2b_bench_c2.json; echo
python scripts/exp_minblocks.py 10000 20000 40000 160000 > gpurun_out/r2b_minblocks.txt 2>&1; cat gpurun_out/r2b_minblocks.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r2b_ncu1.log 2>&1
$CMD > gpurun_out/r2b_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simplyp_quad_kernel -s 7 -c 1 -o gpurun_out/r02_quad_cal_M10000 $CMD > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out | tail -12
