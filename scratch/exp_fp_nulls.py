"""filter->project over NULLABLE columns (10 % NULLs in id and age): NULL-aware two-ring kernel vs the interpreter
kernels (NQE_JIT_NULLS=0).  Algorithmic bytes add the two validity bitmaps (N/8 each) and one output byte-map pass."""
import os, sys
sys.path.insert(0, '.')
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth")
pp = import_module("naive-query-engine_b200.physical_plan")
import bench
ctx = nq.Context(0)
n = int(os.environ.get("N", 100_000_000))
tbl0, bufs = bench.device_table(nq, torch, ctx, synth.FILTER_TABLE, 0, n, [2, 2, 4])
g = torch.Generator(device="cuda"); g.manual_seed(1)
words = (n + 31) // 32
valid = []
nulls = []
for c in range(2):
    w = torch.full((words + 4,), -1, dtype=torch.int32, device="cuda")
    for _ in range(3):  # AND of a few random words -> ~12.5 % zero bits per AND of 3? keep it simple: OR of 3 -> 87.5 % ones
        pass
    r = torch.randint(-2**31, 2**31 - 1, (3, words), dtype=torch.int32, device="cuda", generator=g)
    w[:words] = r[0] | r[1] | r[2]
    valid.append(w)
    nulls.append(n // 8)
tbl = nq.DeviceTable.from_device_pointers(ctx, ["id", "age", "score"], [2, 2, 4], [b.data_ptr() for b in bufs], n,
                                          keepalive=bufs + valid, validity=[valid[0].data_ptr(), valid[1].data_ptr(), 0],
                                          null_counts=nulls + [0])
pred, projs = bench.exprs(nq)
ms = []
for i in range(int(os.environ.get("REPS", 10))):
    out = pp._filter_project(tbl, pred, projs, ["id", "age + 100"])
    ms.append(ctx.last_op_ms)
    rows = out.num_rows
    out.free()
ms = sorted(ms[3:])
knobs = {k: v for k, v in os.environ.items() if k.startswith("NQE_")}
alg = 16 * n + 2 * n / 8 + rows * (16 + 2 / 8)
print("fp-nullable", knobs, "rows", rows, "best %.4f med %.4f ms" % (ms[0], ms[len(ms) // 2]),
      "frac %.3f" % (alg / (ms[len(ms) // 2] * 1e-3) / 1e9 / 6551.4), flush=True)
