#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join" > gpurun_out/pytest_s3c.log 2>&1
tail -3 gpurun_out/pytest_s3c.log
run() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(run X=1; run NQE_JOIN_SIDE=0; run NQE_JOIN_FAT=1; run NQE_JOIN_PART_MB=48; run NQE_JOIN_PART_MB=12) 2>&1 | tee gpurun_out/join_s3c.log
(WHICH=join REPS=2 scratch/launchlist.sh 12 9 python scratch/exp_sec.py; NQE_JOIN_FAT=1 WHICH=join REPS=2 scratch/launchlist.sh 12 9 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3c.log | cut -c1-250
(NQE_HOST_PROF=1 RAW=0 timeout 120 python scratch/exp_e2e.py) 2>&1 | tee gpurun_out/e2e_s3c.log | grep -E "^e2e|pipeline" | tail -3
