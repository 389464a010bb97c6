REPS=5 WHICH=gb,ja python scratch/exp_sec.py 2>&1 | tail -2
python -m pytest tests -m gpu -x -q -k "join or dense or group or key" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ja_probe_scatter -c 1 -s 2 -f -o gpurun_out/ja_probe_full env REPS=3 WHICH=ja python scratch/exp_sec.py > gpurun_out/ja_probe_full.log 2>&1
