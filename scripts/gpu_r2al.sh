#!/bin/bash
# round-2 session AL: the network kernel on config 5 (4096 reaches, 487 levels) with 8 members x 4096 days (2 epochs) under
# ncu application replay
mkdir -p gpurun_out
CMD="python scripts/ncu_network_case.py 5 8 4096"
$CMD > gpurun_out/r2al_plain.log 2>&1 && timeout 900 ncu --replay-mode application --clock-control none -k regex:simplyp_quad_kernel -s 1 -c 1 \
   --section SpeedOfLight --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats \
   --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_global_st.sum,dram__bytes_write.sum,dram__bytes_read.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_active.avg,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum \
   -o gpurun_out/r02b_quad_run_stiff_config5_M8 $CMD > gpurun_out/r2al_ncu.log 2>&1
tail -2 gpurun_out/r2al_ncu.log; tail -1 gpurun_out/r2al_plain.log
