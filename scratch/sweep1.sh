for ps in 0 1 2 3; do echo "== probe shape $ps"; NQE_JA_PROBE_SHAPE=$ps REPS=4 python scratch/exp_ja2.py 2>&1 | tail -3; done
for ss in 1 2; do echo "== split shape $ss (probe 0)"; NQE_PS_SPLIT_SHAPE=$ss REPS=4 python scratch/exp_ja2.py 2>&1 | tail -2; done
for ss in 0 1 2; do echo "== groupby split shape $ss"; NQE_PS_SPLIT_SHAPE=$ss REPS=4 python scratch/exp_gb2.py 2>&1 | tail -2; done
for ps in 1 2 3; do NQE_JA_PROBE_SHAPE=$ps NQE_PS_SPLIT_SHAPE=$((ps%3)) timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "paged or group_by_one_value" 2>&1 | tail -2; done
