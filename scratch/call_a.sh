timeout 600 python -m pytest tests/test_gpu_multi_abi.py tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/exp_multi.py 2>&1 | tail -2
