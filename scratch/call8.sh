#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q -k "join or filter_project or hypothesis or golden" > gpurun_out/pytest_s3h.log 2>&1
tail -3 gpurun_out/pytest_s3h.log
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(runj NQE_JOIN_FUSE=0; runj NQE_JOIN_FUSE=1; runj NQE_JOIN_FUSE=1 NQE_JOIN_SPLIT=2; runj NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=2 NQE_JOIN_GATHER=2) 2>&1 | tee gpurun_out/join_s3h.log
timeout 120 python scratch/exp_fp.py 2>&1 | tail -1 | tee gpurun_out/fp_s3h.log
(WHICH=join REPS=2 scratch/launchlist.sh 12 7 python scratch/exp_sec.py; NQE_JOIN_SPLIT=2 WHICH=join REPS=2 scratch/launchlist.sh 12 7 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3h.log | cut -c1-250
