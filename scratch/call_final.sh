#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3_final.log 2>&1
tail -3 gpurun_out/pytest_s3_final.log
timeout 600 python bench.py > gpurun_out/bench_s3_final.json 2> gpurun_out/bench_s3_final.err
tail -c 600 gpurun_out/bench_s3_final.json; tail -3 gpurun_out/bench_s3_final.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_s3.log 2>&1
tail -2 gpurun_out/bench_under_ncu_s3.log | cut -c1-300
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
