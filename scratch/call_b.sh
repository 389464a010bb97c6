python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_multi_abi.py -m gpu -x -q > gpurun_out/sanitizer_multi_memcheck.log 2>&1
echo "== multi memcheck: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_multi_memcheck.log | tr '\n' ' ')"
python bench.py --gpus 1 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 300 gpurun_out/bench_r02_final.err
python bench.py --impl reference --gpus 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
wc -c gpurun_out/bench_r02_final.json gpurun_out/bench_r02_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
