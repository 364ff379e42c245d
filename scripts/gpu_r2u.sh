#!/bin/bash
# round-2 session U: the network kernel of the final build under ncu application replay (config 3, 64 members x 730 days,
# 3 epochs): store metrics, issue / FP64-pipe activity
mkdir -p gpurun_out
CMD="python scripts/ncu_network_case.py 3 64 730"
$CMD > gpurun_out/r2u_plain.log 2>&1 && ncu --replay-mode application --clock-control none -k regex:simplyp_quad_kernel -s 1 -c 1 \
   --section SpeedOfLight --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis \
   --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_global_st.sum,dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sectors_op_write.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_active.avg \
   -o gpurun_out/r02b_quad_run_stiff_config3_M64 $CMD > gpurun_out/r2u_ncu.log 2>&1
tail -3 gpurun_out/r2u_ncu.log; tail -2 gpurun_out/r2u_plain.log
