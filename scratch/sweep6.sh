for rep in 1 2; do for mb in 48 12 8 6 4; do echo "== slot range $mb MiB"; NQE_JOIN_PART_MB=$mb REPS=5 WHICH=ja python scratch/exp_sec.py 2>&1 | tail -1; done; done
