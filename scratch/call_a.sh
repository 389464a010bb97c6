python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct_table" 2>&1 | tail -5
REPS=5 WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2
for sh in 1 2 3; do NQE_JA_DIRECT_SHAPE=$sh REPS=5 WHICH=ja python scratch/exp_sec.py 2>&1 | tail -1; done
