#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q -k "filter_project or kleene or null or every_operator or error_behaviour or deep or hypothesis" > gpurun_out/pytest_s3b.log 2>&1
tail -3 gpurun_out/pytest_s3b.log
(NQE_JIT_NULLS=1 timeout 120 python scratch/exp_fp_nulls.py) 2>&1 | tee gpurun_out/fp_nulls_s3b.log | grep fp-nullable
(NQE_HOST_PROF=1 RAW=0 timeout 120 python scratch/exp_e2e.py; NQE_HOST_PROF=1 NQE_HOST_CHUNK_ROWS=2097152 RAW=0 timeout 120 python scratch/exp_e2e.py) 2>&1 | tee gpurun_out/e2e_s3b.log | grep -E "^raw|^e2e|pipeline"
REPS=4 timeout 300 bash scratch/launchlist.sh 0 60 python scratch/exp_fp_nulls.py > gpurun_out/launch_fp_nulls_s3b.log 2>&1
tail -16 gpurun_out/launch_fp_nulls_s3b.log
