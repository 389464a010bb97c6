"""group-by on Zipf-1.0 keys (k = floor(G^u)) and on few groups: timing per path"""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import torch, nqe_b200 as nq
from importlib import import_module
synth = import_module("naive-query-engine_b200.synth")
import bench
ctx = nq.Context(0)
n = int(os.environ.get("N", 100_000_000)); G = 100_000
gt, gb = bench.device_table(nq, torch, ctx, synth.GROUPBY_TABLE, 0, n, [2, 4])
u = gb[1].view(torch.float64) / 100.0
zk = torch.clamp(torch.floor(torch.exp(u * float(np.log(G)))).to(torch.int64) - 1, 0, G - 1)
torch.cuda.synchronize()
zt = nq.DeviceTable.from_device_pointers(ctx, ["k", "v"], [2, 4], [zk.data_ptr(), gb[1].data_ptr()], n, keepalive=[zk, gb[1]])
ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(zt.names)
arr = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, 1) for o in range(5)])
for name, t in (("zipf", zt), ("uniform", gt)):
    for i in range(4):
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, t.h, C.pointer(ke), arr, 5, C.byref(h)))
        ms = ctx.last_op_ms
        o = nq.DeviceTable(ctx, h, ["x"] * 5); rows = o.num_rows; o.free()
    print("%s part=%s ms %.3f groups %d" % (name, os.environ.get("NQE_AGG_PART", "default"), ms, rows), flush=True)
