# compute-sanitizer over the direct-table join paths (and the hashed / paged ones behind NQE_JOIN_DIRECT=0) at small sizes
export NQE_JOIN_PART_MIN_ROWS=1000 NQE_JOIN_PART_MIN_MB=0 NQE_AGG_PART_MIN_ROWS=1000 NQE_JOIN_DIRECT_MIN_ROWS=1
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool args..., -- command
  name=$1; shift
  timeout 900 $CS "$@" > gpurun_out/sanitizer_$name.log 2>&1
  echo "== $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$name.log | tail -2 | tr '\n' ' ')"
  grep -E " ok|Error|error:|hazard" gpurun_out/sanitizer_$name.log | sort | uniq -c | sort -rn | head -8
}
WHICH=join run direct_memcheck --tool memcheck --print-limit 20 python scratch/sanitize_run.py
WHICH=join NQE_JOIN_DIRECT_NARROW=0 NQE_JOIN_DIRECT_STAGED=0 run direct_wide_memcheck --tool memcheck --print-limit 20 python scratch/sanitize_run.py
WHICH=join NQE_JOIN_DIRECT=0 run hashed_memcheck --tool memcheck --print-limit 20 python scratch/sanitize_run.py
WHICH=join run direct_racecheck --tool racecheck --print-limit 20 python scratch/sanitize_run.py
