#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q -k "join or distributed or hypothesis" > gpurun_out/pytest_s3e.log 2>&1
tail -3 gpurun_out/pytest_s3e.log
NQE_JOIN_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned_probe_large" 2>&1 | tail -1
run() { timeout 120 env "$@" WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2; echo "   ^ $@"; }
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(run NQE_JOIN_ROWPAY=0 NQE_JOIN_OVERLAP=0 NQE_JOIN_SPLIT=1; runj NQE_JOIN_ROWPAY=1 NQE_JOIN_OVERLAP=0 NQE_JOIN_SPLIT=1; runj NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=1; runj NQE_JOIN_OVERLAP=0 NQE_JOIN_SPLIT=2; run NQE_JOIN_OVERLAP=1 NQE_JOIN_SPLIT=2; run NQE_JOINAGG_PART=1) 2>&1 | tee gpurun_out/join_s3e.log
(NQE_JOIN_OVERLAP=0 WHICH=join,ja REPS=2 scratch/launchlist.sh 12 10 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3e.log | cut -c1-250
