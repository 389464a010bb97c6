#!/bin/bash
run() { timeout 120 env "$@" python scratch/exp_sec.py 2>&1 | tail -3; echo "   ^ $@"; }
run NQE_JOIN_FAT=1
run NQE_JOIN_FAT=0
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_red.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
WHICH=gb REPS=2 timeout 200 ncu --metrics $M --clock-control none -k regex:group_aggregate_kernel -s 1 -c 1 python scratch/exp_sec.py 2>&1 | grep -E "dram|lts|gpu__|sm__|smsp|l1tex"
WHICH=join REPS=2 timeout 200 ncu --metrics $M --clock-control none -k regex:join_probe_kernel -s 1 -c 1 python scratch/exp_sec.py 2>&1 | grep -E "dram|lts|gpu__|sm__|smsp|l1tex"
