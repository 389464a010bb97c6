#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned or hash_join_random or golden_readme" > gpurun_out/pytest_s3i.log 2>&1
tail -2 gpurun_out/pytest_s3i.log
NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=2 NQE_JOIN_GATHER=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned_probe" 2>&1 | tail -1
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(runj NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=3; runj NQE_JOIN_FUSE=1 NQE_JOIN_SPLIT=3; runj NQE_JOIN_FUSE=1 NQE_JOIN_SPLIT=2; runj NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=2 NQE_JOIN_GATHER=2; runj NQE_JOIN_FUSE=1 NQE_JOIN_SPLIT=2 NQE_JOIN_SPLIT_CTAS=4; runj NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=2 NQE_JOIN_GATHER=2 NQE_JOIN_SPLIT_CTAS=4) 2>&1 | tee gpurun_out/join_s3i.log
(NQE_JOIN_FUSE=0 NQE_JOIN_SPLIT=2 NQE_JOIN_GATHER=2 WHICH=join REPS=2 scratch/launchlist.sh 12 8 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3i.log | cut -c1-250
