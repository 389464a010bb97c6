#!/bin/bash
run() { timeout 120 env "$@" WHICH=join,ja python scratch/exp_sec.py 2>&1 | tail -2; echo "   ^ $@"; }
run X=1
run NQE_JOIN_PART=0
run NQE_JOIN_PART=0 NQE_JOIN_FAT=1
run NQE_JOIN_FAT=1 NQE_JOIN_PART_MB=40
run NQE_JOIN_PART_MB=12
