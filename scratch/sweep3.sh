for ps in 0 3 4; do echo "== probe shape $ps"; NQE_JA_PROBE_SHAPE=$ps REPS=4 WHICH=ja python scratch/exp_sec.py 2>&1 | tail -1; done
for ss in 0 3 4; do echo "== split shape $ss"; NQE_PS_SPLIT_SHAPE=$ss REPS=4 WHICH=gb,ja python scratch/exp_sec.py 2>&1 | tail -2; done
echo "== best guess combo"; NQE_PS_SPLIT_SHAPE=3 NQE_JA_PROBE_SHAPE=3 REPS=4 WHICH=gb,ja python scratch/exp_sec.py 2>&1 | tail -2
