"""group-by / join / join+group-by operator timings (device resident), one pass each after warm-up."""
import os, sys, ctypes as C
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n, nb = int(os.environ.get("N", 100_000_000)), int(os.environ.get("NB", 10_000_000))
reps = int(os.environ.get("REPS", 3))
which = os.environ.get("WHICH", "gb,join,ja").split(",")
I64, F64 = 2, 4
col = nq.ColumnExpr.try_create
AGG5 = [(0, None), (1, None), (2, None), (3, None), (4, None)]
def aggs(c): return (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, c) for o, _ in AGG5])
if "gb" in which:
    gt, gb = bench.device_table(nq, torch, ctx, synth.GROUPBY_TABLE, 0, n, [I64, F64])
    ke, keep = col(None, 0).to_expr(gt.names)
    for i in range(reps):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, gt.h, C.pointer(ke), aggs(1), 5, C.byref(h)))
        ms = ctx.last_op_ms
        t = nq.DeviceTable(ctx, h, ["x"] * 5); rows = t.num_rows; t.free()
    print("group_by ms %.3f groups %d frac %.3f" % (ms, rows, 16.0 * n / (ms * 1e-3) / 1e9 / 6551.4), flush=True)
    gt.free(); del gb; torch.cuda.empty_cache()
if "join" in which or "ja" in which:
    lt0, lb0 = bench.device_table(nq, torch, ctx, synth.join_build_table(nb), 0, nb, [I64])
    lt = pp._filter_project(lt0, None, [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 0), "Modulos", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(100000)))], ["k", "a"])
    rt, rb = bench.device_table(nq, torch, ctx, synth.join_probe_table(nb), 0, n, [I64, F64])
if "join" in which:
    for i in range(reps):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_join(ctx.h, lt.h, rt.h, 0, 0, C.byref(h)))
        ms = ctx.last_op_ms
        t = nq.DeviceTable(ctx, h, ["k", "a", "fk", "b"]); rows = t.num_rows; t.free()
    print("hash_join ms %.3f rows %d frac %.3f" % (ms, rows, (16.0 * nb + 48.0 * n) / (ms * 1e-3) / 1e9 / 6551.4), flush=True)
if "ja" in which:
    for i in range(reps):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, 1, aggs(3), 5, C.byref(h)))
        ms = ctx.last_op_ms
        t = nq.DeviceTable(ctx, h, ["x"] * 5); rows = t.num_rows; t.free()
    print("join_group_by ms %.3f groups %d frac %.3f" % (ms, rows, (16.0 * nb + 16.0 * n) / (ms * 1e-3) / 1e9 / 6551.4), flush=True)
