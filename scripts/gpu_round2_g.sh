#!/bin/bash
# round-2 GPU session G: full capture of the large-ensemble build (1.6x10^5 members), then the final bench lines
mkdir -p gpurun_out
CMD="python bench.py --members 160000 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simplyp_quad_kernel -s 7 -c 1 -o gpurun_out/r02_quad_cal_M160000 $CMD > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err; tail -c 200 gpurun_out/r2g_bench_c2.err; head -c 330 gpurun_out/r2g_bench_c2.json; echo
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; head -c 330 gpurun_out/r2g_bench_c4.json; echo
python bench.py --period full --members 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_full_period.json 2> gpurun_out/r2g_bench_full_period.err; head -c 330 gpurun_out/r2g_bench_full_period.json; echo
