python bench.py --gpus 1 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 600 gpurun_out/bench_r02_final.err
python bench.py --impl reference --gpus 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
wc -c gpurun_out/bench_r02_final.json gpurun_out/bench_r02_reference_arm.json
