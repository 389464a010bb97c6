# pipelined direct probe+scatter of the fused join -> group-by: NQE_JA_DIRECT_PIPE 0 = plain, 1..5 = shapes
for pp in 0 1 2 3 4 5; do echo "== pipe $pp"; NQE_JA_DIRECT_PIPE=$pp REPS=5 WHICH=ja timeout 300 python scratch/exp_sec.py 2>&1 | tail -1; done
NQE_JA_DIRECT_PIPE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join_aggregate" 2>&1 | tail -2
