#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter_project_random or nullable or kleene or null_predicate or join_aggregate or hash_join_random or partitioned" > gpurun_out/pytest_s3d.log 2>&1
tail -3 gpurun_out/pytest_s3d.log
run() { timeout 120 env "$@" WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2; echo "   ^ $@"; }
(run NQE_JOIN_L2=0; run NQE_JOIN_L2=1; run NQE_JOIN_L2=3) 2>&1 | tee gpurun_out/join_s3d.log
(NQE_JIT_NULLS=1 timeout 120 python scratch/exp_fp_nulls.py) 2>&1 | tee gpurun_out/fp_nulls_s3d.log | grep fp-nullable
(WHICH=join REPS=2 scratch/launchlist.sh 12 9 python scratch/exp_sec.py) 2>&1 | tee gpurun_out/join_launch_s3d.log | cut -c1-250
