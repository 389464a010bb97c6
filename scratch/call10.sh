#!/bin/bash
mkdir -p gpurun_out
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(runj X=1; runj NQE_JOIN_PART_MB=12; runj NQE_JOIN_PART_MB=48) 2>&1 | tee gpurun_out/join_s3j.log
(NQE_JOIN_PART_MB=12 WHICH=join REPS=2 scratch/launchlist.sh 12 8 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3j.log | cut -c1-250
