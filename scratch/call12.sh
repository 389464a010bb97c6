#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join" > gpurun_out/pytest_s3l.log 2>&1
tail -2 gpurun_out/pytest_s3l.log
timeout 300 python bench.py > gpurun_out/bench_s3_final2.json 2> gpurun_out/bench_s3_final2.err
tail -c 300 gpurun_out/bench_s3_final2.json; tail -2 gpurun_out/bench_s3_final2.err
