#!/bin/bash
run() { timeout 60 env "$@" REPS=10 python scratch/exp_fp.py 2>&1 | tail -1; }
run
run NQE_JIT_TMA_LAG=2
run NQE_JIT_TMA_LAG=3
run NQE_JIT_TMA_LAG=5
run NQE_JIT_TMA_LAG=3 NQE_JIT_TMA_WALKERS=1
