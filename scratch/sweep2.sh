export NQE_PS_SPLIT_SHAPE=1
for ps in 0 2 3; do echo "== probe shape $ps"; NQE_JA_PROBE_SHAPE=$ps REPS=4 WHICH=ja python scratch/exp_sec.py 2>&1 | tail -1; done
echo "== all, paged"; NQE_JA_PROBE_SHAPE=2 REPS=4 python scratch/exp_sec.py 2>&1 | tail -3
echo "== all, direct fused + table group-by"; NQE_JOINAGG_PAGED=0 NQE_AGG_PART=0 REPS=4 python scratch/exp_sec.py 2>&1 | tail -3
NQE_JA_PROBE_SHAPE=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join" 2>&1 | tail -2
