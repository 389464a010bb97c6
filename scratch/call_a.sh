# final launch lists of round 2 (metrics pass) for group-by, hash join, join->group-by; plus a full capture of the direct-join kernels
for w in gb join ja; do
  REPS=1 WHICH=$w bash scratch/launchlist.sh 0 80 python scratch/exp_sec.py > gpurun_out/launchlist_final_$w.txt 2>&1
  cp /tmp/ll.csv gpurun_out/launchlist_final_$w.csv
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"join_direct|ja_direct_scatter|gp2_aggregate|ps_split" -c 6 -f -o gpurun_out/final_ops_full env REPS=1 WHICH=gb,join,ja python scratch/exp_sec.py > gpurun_out/final_ops_full.log 2>&1
tail -2 gpurun_out/final_ops_full.log
