#!/bin/bash
run() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
run X=1
run NQE_JOIN_FAT=0
run NQE_JOIN_PART=0
run NQE_JOIN_PART_MB=12
run NQE_JOIN_PART_MB=40
WHICH=join REPS=2 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s 30 -c 40 python scratch/exp_sec.py 2>&1 | grep -E "^  [a-z<v].*\(|gpu__time|dram__|lts__" | paste - - - - - | awk '{print $1, $2, $(NF-13), $(NF-10), $(NF-9), $(NF-6), $(NF-5), $(NF-2), $(NF-1), $NF}' | tail -12
