timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct_table or hash_join" 2>&1 | tail -3
REPS=5 WHICH=join timeout 300 python scratch/exp_sec.py 2>&1 | tail -1
NB=5000000 REPS=5 WHICH=join timeout 300 python scratch/exp_sec.py 2>&1 | tail -1
REPS=1 WHICH=join bash scratch/launchlist.sh 0 60 python scratch/exp_sec.py | tail -4
