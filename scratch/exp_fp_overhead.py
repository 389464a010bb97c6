"""filter->project step time: through the Python host mirror vs bare C-ABI calls (prebuilt expression structs)"""
import sys, os, time, ctypes as C
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
n = 100_000_000
with torch.cuda.stream(stream):
    tbl, bufs = bench.device_table(nq, torch, ctx, synth.FILTER_TABLE, 0, n, [2, 2, 4])
pred, projs = bench.exprs(nq)
names = ["id", "age + 100"]
def step_py():
    out = pp._filter_project(tbl, pred, projs, names); out.free()
pe, k1 = pred.to_expr(tbl.names)
earr = (nq._ffi.Expr * 2)()
keep = []
for i, e in enumerate(projs):
    ex, arr = e.to_expr(tbl.names); keep.append(arr); earr[i] = ex
lib = ctx.lib
def step_c():
    h = C.c_void_p()
    rc = lib.nqe_filter_project(ctx.h, tbl.h, C.byref(pe), earr, 2, C.byref(h))
    assert rc == 0
    lib.nqe_table_free(h)
for name, f in (("python mirror", step_py), ("bare C ABI", step_c)):
    for _ in range(5): f()
    kms = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(50):
        f(); kms.append(ctx.last_op_ms)
    e1.record(stream); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    print("%-14s step %.4f ms (events) %.4f ms (wall) kernel %.4f ms" % (name, e0.elapsed_time(e1) / 50, wall, sum(kms) / 50), flush=True)
