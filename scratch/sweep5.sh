for t in 0 1 2 3; do echo "== split tma $t"; NQE_PS_SPLIT_TMA=$t REPS=4 WHICH=gb,ja python scratch/exp_sec.py 2>&1 | tail -2; done
echo "== hashed keys (NQE_AGG_DENSE=0)"; for t in 0 1 3; do NQE_AGG_DENSE=0 NQE_PS_SPLIT_TMA=$t REPS=4 WHICH=gb python scratch/exp_sec.py 2>&1 | tail -1; done
NQE_PS_SPLIT_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_by or group_key or paged" 2>&1 | tail -2
