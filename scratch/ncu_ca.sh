#!/bin/bash
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for cfg in "NQE_JIT_IMPL=ca NQE_JIT_D=4" "NQE_JIT_IMPL=ca NQE_JIT_D=2" "NQE_JIT_IMPL=ca NQE_JIT_D=2 NQE_JIT_L2_HINTS=1" "NQE_JIT_IMPL=ca NQE_JIT_D=1" "NQE_JIT_TMA_K=8 NQE_JIT_TMA_S=3"; do
  echo "=== $cfg"
  env $cfg REPS=5 timeout 120 ncu --metrics $M --clock-control none -k regex:nqe_fp_jit -s 3 -c 1 python scratch/exp_fp.py 2>&1 | grep -E "dram__|lts__|gpu__time|sm__warps|fp \{" 
done
