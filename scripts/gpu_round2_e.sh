#!/bin/bash
# round-2 GPU session E: bench lines of configs 3 and 5 on the final build (+ reference arms), store metrics of the
# network kernel by ncu application replay
mkdir -p gpurun_out
python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err; tail -c 300 gpurun_out/r2e_bench_c3.err; head -c 330 gpurun_out/r2e_bench_c3.json; echo
python bench.py --config 3 --members 256 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 8 > gpurun_out/r2e_bench_c3_m256.json 2> gpurun_out/r2e_bench_c3_m256.err; head -c 330 gpurun_out/r2e_bench_c3_m256.json; echo
python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err; tail -c 300 gpurun_out/r2e_bench_c5.err; head -c 330 gpurun_out/r2e_bench_c5.json; echo
python bench.py --config 5 --members 8 --steps 2 --warmup 3 --no-cpu-baseline --e2e-members 1 > gpurun_out/r2e_bench_c5_m8.json 2> gpurun_out/r2e_bench_c5_m8.err; head -c 330 gpurun_out/r2e_bench_c5_m8.json; echo
python bench.py --config 5 --members 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c5_m1.json 2> gpurun_out/r2e_bench_c5_m1.err; head -c 330 gpurun_out/r2e_bench_c5_m1.json; echo
for c in 2 3 5; do python bench.py --impl reference --config $c --steps 3 --warmup 1 > gpurun_out/r2e_ref_c$c.json 2> gpurun_out/r2e_ref_c$c.err; head -c 250 gpurun_out/r2e_ref_c$c.json; echo; done
CMD="python scripts/ncu_network_case.py 3 8 730"
$CMD > gpurun_out/r2e_plain.log 2>&1 && ncu --replay-mode application --clock-control none -k regex:simplyp_quad_kernel -s 1 -c 1 \
   --section SpeedOfLight --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis \
   --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_global_st.sum,dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sectors_op_write.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
   -o gpurun_out/r02_quad_run_stiff_config3_M8 $CMD > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
