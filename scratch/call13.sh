#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_by_one_value" 2>&1 | tail -4
NQE_AGG_PART=1 NQE_AGG_PART_MIN_ROWS=1000 timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_by" 2>&1 | tail -6
rung() { timeout -s KILL 90 env "$@" WHICH=gb python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
(rung NQE_AGG_PART=0; rung NQE_AGG_PART=1) 2>&1 | tee gpurun_out/gb_s3m.log
