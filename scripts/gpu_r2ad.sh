#!/bin/bash
# round-2 session AD: captures of the two other builds of the final library — the 128-register ensemble build at
# 1.6x10^5 members (kernel replay, --set full) and the 168-register network build on config 3 at 256 members x 730 days
# (application replay)
mkdir -p gpurun_out
CMD="python bench.py --members 160000 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2ad_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simplyp_quad_kernel -s 7 -c 1 -o gpurun_out/r02b_quad_cal_M160000 $CMD > gpurun_out/r2ad_ncu1.log 2>&1
tail -2 gpurun_out/r2ad_ncu1.log
CMD="python scripts/ncu_network_case.py 3 256 730"
$CMD > gpurun_out/r2ad_plain2.log 2>&1 && ncu --replay-mode application --clock-control none -k regex:simplyp_quad_kernel -s 1 -c 1 \
   --section SpeedOfLight --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis \
   --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_global_st.sum,dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sectors_op_write.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_active.avg,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum \
   -o gpurun_out/r02b_quad_run_stiff_config3_M256 $CMD > gpurun_out/r2ad_ncu2.log 2>&1
tail -2 gpurun_out/r2ad_ncu2.log; tail -1 gpurun_out/r2ad_plain2.log
