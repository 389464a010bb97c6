"""fused join -> group-by (configs[3] composite) timing + checksum; NQE_JOINAGG_PAGED=0/1"""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n, nb = int(os.environ.get("N", 100_000_000)), 10_000_000
I64, F64 = 2, 4
col = nq.ColumnExpr.try_create
lt0, lb0 = bench.device_table(nq, torch, ctx, synth.join_build_table(nb), 0, nb, [I64])
lt = pp._filter_project(lt0, None, [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 0), "Modulos", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(100000)))], ["k", "a"])
rt, rb = bench.device_table(nq, torch, ctx, synth.join_probe_table(nb), 0, n, [I64, F64])
aggs = [(0, 3), (1, 3), (2, 3), (3, 3), (4, 3)]
arr = (nq._ffi.Agg * len(aggs))(*[nq._ffi.Agg(o, c) for o, c in aggs])
for i in range(int(os.environ.get("REPS", 6))):
    h = C.c_void_p()
    ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, 1, arr, len(aggs), C.byref(h)))
    ms = ctx.last_op_ms
    t = nq.DeviceTable(ctx, h, ["count", "sum", "avg", "min", "max"])
    if i == 0:
        out = t.to_arrow()
    rows = t.num_rows
    t.free()
    print("joinagg paged=%s ms %.3f groups %d" % (os.environ.get("NQE_JOINAGG_PAGED", "default"), ms, rows), flush=True)
c = out.column(0).to_numpy(); s = out.column(1).to_numpy(); mn = out.column(3).to_numpy(); mx = out.column(4).to_numpy()
o = np.lexsort((mx, mn, c))
print("checksum rows=%d count_sum=%d sum_sum=%.6f min_min=%r max_max=%r sorted_sum_dot=%.6f" %
      (len(c), int(c.sum()), float(np.sort(s).sum()), float(mn.min()), float(mx.max()),
       float((s[o] * np.arange(len(o))).sum())), flush=True)
