# compute-sanitizer over every operator path at small sizes; summaries -> gpurun_out/sanitizer_*.log
export NQE_JOIN_PART_MIN_ROWS=1000 NQE_JOIN_PART_MIN_MB=0 NQE_AGG_PART_MIN_ROWS=1000
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool args..., -- command
  name=$1; shift
  timeout 1500 $CS "$@" > gpurun_out/sanitizer_$name.log 2>&1
  echo "== $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$name.log | tail -2 | tr '\n' ' ')"
  grep -E "all ok|Error|error:|hazard" gpurun_out/sanitizer_$name.log | sort | uniq -c | sort -rn | head -8
}
run memcheck --tool memcheck --print-limit 20 python scratch/sanitize_run.py
NQE_JIT_IMPL=ca WHICH=fp run memcheck_ca --tool memcheck --print-limit 20 python scratch/sanitize_run.py
NQE_JIT=0 WHICH=fp run memcheck_interp --tool memcheck --print-limit 20 python scratch/sanitize_run.py
NQE_JOINAGG_PAGED=0 NQE_AGG_PART=0 WHICH=join run memcheck_direct --tool memcheck --print-limit 20 python scratch/sanitize_run.py
run racecheck --tool racecheck --print-limit 20 python scratch/sanitize_run.py
