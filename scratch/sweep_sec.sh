#!/bin/bash
run() { timeout 120 env "$@" WHICH=join python scratch/exp_sec.py 2>&1 | tail -1; echo "   ^ $@"; }
run X=1
run NQE_JOIN_PROBE_K=8
WHICH=join REPS=2 scratch/launchlist.sh 12 9 python scratch/exp_sec.py 2>&1 | tail -9
