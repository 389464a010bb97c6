"""group-by cost per aggregate combination (which part of the L2 work dominates?)"""
import sys, ctypes as C
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth")
import bench
ctx = nq.Context(0)
n = 100_000_000
gt, gb = bench.device_table(nq, torch, ctx, synth.GROUPBY_TABLE, 0, n, [2, 4])
ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(gt.names)
NAMES = ["count", "sum", "avg", "min", "max"]
for ops in [[0], [1], [3], [4], [0, 1], [3, 4], [0, 1, 3], [0, 1, 2, 3, 4]]:
    arr = (nq._ffi.Agg * len(ops))(*[nq._ffi.Agg(o, 1) for o in ops])
    for i in range(3):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, gt.h, C.pointer(ke), arr, len(ops), C.byref(h)))
        ms = ctx.last_op_ms
        t = nq.DeviceTable(ctx, h, ["x"] * len(ops)); rows = t.num_rows; t.free()
    print("groupby", [NAMES[o] for o in ops], "ms", round(ms, 3), flush=True)
