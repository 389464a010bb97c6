"""group-by (configs[2]) timing + a path-independent checksum of the result; run once per knob setting:
   NQE_AGG_PART=0 python scratch/exp_gb2.py ; NQE_AGG_PART=1 python scratch/exp_gb2.py"""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth")
import bench
ctx = nq.Context(0)
n = int(os.environ.get("N", 100_000_000))
gt, gb = bench.device_table(nq, torch, ctx, synth.GROUPBY_TABLE, 0, n, [2, 4])
ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(gt.names)
ops = [0, 1, 2, 3, 4]
arr = (nq._ffi.Agg * len(ops))(*[nq._ffi.Agg(o, 1) for o in ops])
for i in range(int(os.environ.get("REPS", 6))):
    h = C.c_void_p()
    ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, gt.h, C.pointer(ke), arr, len(ops), C.byref(h)))
    ms = ctx.last_op_ms
    t = nq.DeviceTable(ctx, h, ["count", "sum", "avg", "min", "max"])
    if i == 0:
        out = t.to_arrow()
    rows = t.num_rows
    t.free()
    print("groupby part=%s ms %.3f groups %d" % (os.environ.get("NQE_AGG_PART", "default"), ms, rows), flush=True)
c = out.column(0).to_numpy(); s = out.column(1).to_numpy(); mn = out.column(3).to_numpy(); mx = out.column(4).to_numpy()
o = np.lexsort((mx, mn, c))
print("checksum rows=%d count_sum=%d sum_sum=%.6f min_min=%r max_max=%r sorted_sum_dot=%.6f" %
      (len(c), int(c.sum()), float(np.sort(s).sum()), float(mn.min()), float(mx.max()),
       float((s[o] * np.arange(len(o))).sum())), flush=True)
