"""end-to-end filter->project (pinned host in, pinned host out) for the current NQE_HOST_* knobs, next to the raw
PCIe copy times of the same bytes (H2D alone, D2H alone, both directions at once)."""
import os, sys, time
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
host = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(3)]
for h, b in zip(host, bufs):
    h.copy_(b)
res = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(2)]
torch.cuda.synchronize()
def ev():
    return torch.cuda.Event(enable_timing=True)
if os.environ.get("RAW", "1") == "1":
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name in ("h2d", "d2h", "both"):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    bufs[0].copy_(host[0], non_blocking=True); bufs[1].copy_(host[1], non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    res[0][: n // 2].copy_(bufs[2][: n // 2], non_blocking=True); res[1][: n // 2].copy_(bufs[2][n // 2:], non_blocking=True)
            torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
        print("raw", name, "%.2f ms" % best, flush=True)
ms = []
for i in range(int(os.environ.get("REPS", 6))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rows, _ = pp.filter_project_host(ctx, ["id", "age", "score"], [2, 2, 4], [h.data_ptr() for h in host], n, pred, projs,
                                     [r.data_ptr() for r in res], n)
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
ms = sorted(ms[1:])
print("e2e", {k: v for k, v in os.environ.items() if k.startswith("NQE_")}, "rows", rows, "best %.2f med %.2f ms" % (ms[0], ms[len(ms) // 2]), flush=True)
