"""filter->project kernel time over selectivities (id < K, ids uniform in [0, 1000)); knob NQE_JIT_SPARSE_DIV in the environment"""
import os, sys
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth"); pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n = 100_000_000
ft, fb = bench.device_table(nq, torch, ctx, synth.FILTER_TABLE, 0, n, [2, 2, 4])
col, lit, sv = nq.ColumnExpr.try_create, nq.PhysicalLiteralExpr.create, nq.ScalarValue
out_line = []
for kk in (0, 1, 10, 50, 100, 200, 500, 900):
    pred = nq.PhysicalBinaryExpr.create(col(None, 0), "Lt", lit(sv.Int64(kk)))
    projs = [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 1), "Plus", lit(sv.Int64(100)))]
    kms = []
    for _ in range(8):
        out = pp._filter_project(ft, pred, projs, ["id", "age + 100"]); kms.append(ctx.last_op_ms); rows = out.num_rows; out.free()
    k = sorted(kms[2:])[3]
    out_line.append("s=%.3f %.3f ms" % (rows / n, k))
print("sparse_div", os.environ.get("NQE_JIT_SPARSE_DIV", "default"), " | ".join(out_line), flush=True)
