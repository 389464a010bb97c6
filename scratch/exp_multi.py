"""single-process multi-GPU broadcast-build join -> group-by through the C ABI (nqe_multi_join_aggregate): wall time per call.
MEMBERS=n devices (default: all), 1.25e8 probe rows per member, 1e7 build rows on member 0."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
nd = int(os.environ.get("MEMBERS", torch.cuda.device_count()))
per = int(os.environ.get("PER", 125_000_000)); nb = 10_000_000
m = nq.MultiContext(list(range(nd)))
col = nq.ColumnExpr.try_create
shards, keep = [], []
for i, c in enumerate(m.members):
    torch.cuda.set_device(i)
    t, b = bench.device_table(nq, torch, c, synth.join_probe_table(nb), i * per, per, [2, 4])
    shards.append(t); keep.append(b)
torch.cuda.set_device(0)
lt0, lb0 = bench.device_table(nq, torch, m.members[0], synth.join_build_table(nb), 0, nb, [2])
left = pp._filter_project(lt0, None, [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 0), "Modulos", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(100000)))], ["k", "a"])
aggs = [(5, 0), (0, 3), (1, 3), (2, 3), (3, 3), (4, 3)]
ts = []
for r in range(6):
    t0 = time.perf_counter()
    out = m.join_aggregate(left, shards, 0, 0, 1, aggs, ["key", "count", "sum", "avg", "min", "max"])
    ts.append((time.perf_counter() - t0) * 1e3)
    rows = out.num_rows
    if r == 5:
        tab = out.to_arrow()
        cnt = int(np.asarray(tab.column(1).to_numpy()).sum())
    out.free()
print("members %d probe rows %d groups %d joined rows %d  ms per call: %s  best %.2f  -> %.3e probe rows/s" %
      (nd, nd * per, rows, cnt, " ".join("%.2f" % t for t in ts), min(ts), nd * per / (min(ts) / 1e3)), flush=True)
