# how much does the direct table's L2 residency matter: build sides of 2.5e6 / 5e6 / 7.5e6 / 1e7 keys = 20 / 40 / 60 / 80 MB tables
for nb in 2500000 5000000 7500000 10000000; do echo "== build rows $nb"; NB=$nb REPS=5 WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2; done
