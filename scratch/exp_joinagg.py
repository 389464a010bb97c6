import sys, time, ctypes as C
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n, nb = 100_000_000, 10_000_000
I64, F64 = 2, 4
col = nq.ColumnExpr.try_create
lt0, lb0 = bench.device_table(nq, torch, ctx, synth.join_build_table(nb), 0, nb, [I64])
lt = pp._filter_project(lt0, None, [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 0), "Modulos", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(100000)))], ["k", "a"])
rt, rb = bench.device_table(nq, torch, ctx, synth.join_probe_table(nb), 0, n, [I64, F64])
def run(aggs, group=1):
    arr = (nq._ffi.Agg * len(aggs))(*[nq._ffi.Agg(o, c) for o, c in aggs])
    for i in range(3):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, group, arr, len(aggs), C.byref(h)))
        ms = ctx.last_op_ms
        t = nq.DeviceTable(ctx, h, ["x"] * len(aggs)); rows = t.num_rows; t.free()
    print(aggs, "group", group, "ms", round(ms, 3), "groups", rows, flush=True)
run([(0, 3)]); run([(1, 3)]); run([(3, 3)]); run([(4, 3)]); run([(0, 3), (1, 3)]); run([(0,3),(1,3),(2,3),(3,3),(4,3)])
run([(0, 3)], group=2)  # group by fk (1e7 groups, probe side)
