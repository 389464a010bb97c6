#!/bin/bash
mkdir -p gpurun_out
NQE_AGG_PART=1 NQE_AGG_PART_MIN_ROWS=1000 timeout -s KILL 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_by_one_value and 300000" 2>&1 | tail -2
(NQE_AGG_PART=1 timeout -s KILL 60 env WHICH=gb python scratch/exp_sec.py 2>&1 | tail -1) | tee gpurun_out/gb_s3o.log
