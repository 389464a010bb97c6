#!/bin/bash
# usage: launchlist.sh <skip> <count> <cmd...>   -> per-kernel time / dram bytes / L2 hit of the captured launches
S=$1; C=$2; shift 2
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s $S -c $C --csv --log-file /tmp/ll.csv "$@" > /tmp/ll.out 2>&1 || tail -5 /tmp/ll.out
python - <<'PY'
import csv, collections
rows = [l for l in open('/tmp/ll.csv') if not l.startswith('==')]
agg = collections.OrderedDict()
for x in csv.DictReader(rows):
    k = (x['ID'], x['Kernel Name'][:60])
    agg.setdefault(k, {})[x['Metric Name']] = (x['Metric Value'], x['Metric Unit'])
for (i, k), m in agg.items():
    print(i, k, ' | '.join(f"{n.split('__')[1][:18]}={v[0]}{v[1]}" for n, v in m.items()))
PY
