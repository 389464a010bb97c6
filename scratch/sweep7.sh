for dp in 148 74 49 37; do echo "== dense parts $dp"; NQE_DENSE_PARTS=$dp REPS=5 WHICH=gb,ja python scratch/exp_sec.py 2>&1 | tail -2; done
