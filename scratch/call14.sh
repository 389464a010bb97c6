#!/bin/bash
mkdir -p gpurun_out
(NQE_AGG_PART=1 WHICH=gb REPS=2 scratch/launchlist.sh 0 30 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/gb_launch_s3n.log | cut -c1-250 | tail -14
