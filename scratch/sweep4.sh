export NQE_JA_PROBE_SHAPE=4
for mb in 48 24 12 10; do echo "== slot range $mb MiB"; NQE_JOIN_PART_MB=$mb REPS=4 WHICH=ja python scratch/exp_sec.py 2>&1 | tail -1; done
