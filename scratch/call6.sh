#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join" > gpurun_out/pytest_s3f.log 2>&1
tail -3 gpurun_out/pytest_s3f.log
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(runj NQE_JOIN_OVERLAP=0 NQE_JOIN_SPLIT=1; runj NQE_JOIN_OVERLAP=0 NQE_JOIN_SPLIT=2; runj NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=2; runj NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=2 NQE_JOIN_SPLIT_CTAS=4; runj NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=2 NQE_JOIN_SPLIT_CTAS=5;  runj NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=1 NQE_JOIN_SPLIT_CTAS=4) 2>&1 | tee gpurun_out/join_s3f.log
(NQE_JOIN_OVERLAP=0 WHICH=join REPS=2 scratch/launchlist.sh 12 7 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3f.log | cut -c1-250
