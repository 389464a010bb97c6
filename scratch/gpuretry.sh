#!/bin/bash
# usage: gpuretry.sh <timeout> <script>   -- retries while the pod answers busy/transient (nothing is charged for those)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box"; then sleep 90; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up after 20 tries"
