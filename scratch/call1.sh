#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -m gpu -x -q -k "filter_project or kleene or null or every_operator or error_behaviour or deep or hypothesis or golden" > gpurun_out/pytest_s3a.log 2>&1
tail -3 gpurun_out/pytest_s3a.log
(NQE_JIT_NULLS=0 timeout 120 python scratch/exp_fp_nulls.py; NQE_JIT_NULLS=1 timeout 120 python scratch/exp_fp_nulls.py) 2>&1 | tee gpurun_out/fp_nulls_s3a.log | grep fp-nullable
(timeout 120 python scratch/exp_e2e.py; for c in 2097152 4194304 16777216; do RAW=0 NQE_HOST_CHUNK_ROWS=$c timeout 120 python scratch/exp_e2e.py; done) 2>&1 | tee gpurun_out/e2e_s3a.log | grep -E "^raw|^e2e"
