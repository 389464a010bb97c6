timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct_table or hash_join or join_aggregate" 2>&1 | tail -3
REPS=5 WHICH=join,ja timeout 300 python scratch/exp_sec.py 2>&1 | tail -2
REPS=1 WHICH=join,ja bash scratch/launchlist.sh 0 60 python scratch/exp_sec.py | grep -v synth | cut -c1-50,60-300 | tail -16
