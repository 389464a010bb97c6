"""filter->project kernel timing for the current env knobs (one process per configuration)."""
import os, sys
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth")
pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n = int(os.environ.get("N", 100_000_000))
tbl, bufs = bench.device_table(nq, torch, ctx, synth.FILTER_TABLE, 0, n, [2, 2, 4])
pred, projs = bench.exprs(nq)
ms = []
for i in range(int(os.environ.get("REPS", 12))):
    out = pp._filter_project(tbl, pred, projs, ["id", "age + 100"])
    ms.append(ctx.last_op_ms)
    rows = out.num_rows
    out.free()
ms = sorted(ms[3:])
knobs = {k: v for k, v in os.environ.items() if k.startswith("NQE_")}
print("fp", knobs, "rows", rows, "best %.4f med %.4f ms" % (ms[0], ms[len(ms) // 2]), "frac %.3f" % ((16 * n + 16 * rows) / (ms[len(ms)//2] * 1e-3) / 1e9 / 6551.4), flush=True)
