#!/bin/bash
mkdir -p gpurun_out
runj() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(runj NQE_JOIN_PART_MB=96; runj NQE_JOIN_PART_MB=48) 2>&1 | tee gpurun_out/join_s3k.log
timeout 600 python -m pytest tests/test_gpu_alternate_paths.py -m gpu -x -q > gpurun_out/pytest_s3k.log 2>&1
tail -15 gpurun_out/pytest_s3k.log
