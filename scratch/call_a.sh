timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct_table or hash_join or join_aggregate or golden_readme" 2>&1 | tail -2
REPS=7 WHICH=join,ja timeout 300 python scratch/exp_sec.py 2>&1 | tail -2
