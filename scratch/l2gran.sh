#!/bin/bash
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:two_sectors --csv --log-file /tmp/g.csv scratch/l2gran | head -1
python - <<'PY'
import csv
rows=[l for l in open('/tmp/g.csv') if not l.startswith('==')]
agg={}
for x in csv.DictReader(rows):
    agg.setdefault(x['ID'],{'k':x['Kernel Name'][:40]})[x['Metric Name']]=x['Metric Value']+x['Metric Unit']
for i,m in agg.items(): print(i, m)
PY
