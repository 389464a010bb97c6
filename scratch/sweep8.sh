# L2 policies of the direct-table probes: bit 0 = streams evict_first, bit 1 = table reads evict_last
for cm in 0 1 2 3; do echo "== cache mode $cm"; NQE_JOIN_DIRECT_CACHE=$cm NQE_JA_DIRECT_CACHE=$cm REPS=5 WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2; done
