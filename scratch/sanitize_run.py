"""Every operator path at small sizes, for compute-sanitizer (memcheck / racecheck):
    NQE_JOIN_PART_MIN_ROWS=1000 NQE_JOIN_PART_MIN_MB=0 NQE_AGG_PART_MIN_ROWS=1000 \
        compute-sanitizer --tool memcheck python scratch/sanitize_run.py
The thresholds are lowered so that the partitioned probe, the paged fused join -> group-by and the paged shared-memory
group-by run on inputs a sanitizer can afford; results are checked against the oracle as in smoke()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyarrow as pa

import __graft_entry__ as ge
import nqe_b200 as nq
from oracle import oracle as O

rng = np.random.default_rng(1)
col, lit, sv = nq.ColumnExpr.try_create, nq.PhysicalLiteralExpr.create, nq.ScalarValue


def fp(n, nulls):
    ids = rng.integers(0, 1000, n)
    age = rng.integers(0, 100, n)
    mask = rng.random(n) < 0.1 if nulls else None
    rb = pa.RecordBatch.from_arrays([pa.array(ids, mask=mask), pa.array(age)], names=["id", "age"])
    scan = nq.ScanPlan.create(nq.MemTable.try_create(rb.schema, [rb]), None)
    pred = nq.PhysicalBinaryExpr.create(col(None, 0), "Lt", lit(sv.Int64(500)))
    plan = nq.ProjectionPlan.create(nq.SelectionPlan.create(scan, pred), pa.schema([("id", pa.int64()), ("a", pa.int64())]),
                                    [col(None, 0), nq.PhysicalBinaryExpr.create(col(None, 1), "Plus", lit(sv.Int64(100)))])
    got = plan.execute()[0]
    ob = O.Batch(["id", "age"], [O.Col("i64", ids, None if mask is None else (~mask).astype(np.uint8)), O.Col("i64", age)])
    want = O.projection(O.selection(ob, ("bin", "Lt", ("col", 0), ("lit", "i64", 500))),
                        [("col", 0), ("bin", "Plus", ("col", 1), ("lit", "i64", 100))])
    assert got.num_rows == want.num_rows
    assert got.column(0).to_pylist() == want.cols[0].to_pylist() and got.column(1).to_pylist() == want.cols[1].to_pylist()
    print("filter_project", n, "nulls" if nulls else "", "ok", flush=True)


which = os.environ.get("WHICH", "fp,join,agg,misc").split(",")
if "fp" in which:
    for n in (0, 1, 2049, 70_001):
        fp(n, False)
    fp(70_001, True)
if "join" in which or "agg" in which:
    # direct kernels, then (with the lowered thresholds) the partitioned probe / paged paths
    for nl, nr, groups in ((500, 6000, 37), (20_000, 150_000, 5000)):
        lk = rng.permutation(nl).astype(np.int64) * 3 + 1
        la = rng.integers(0, groups, nl).astype(np.int64)
        fk = lk[rng.integers(0, nl, nr)]
        fk[rng.random(nr) < 0.2] += 1
        b = np.round(rng.normal(0, 10, nr), 3)
        ge.join_and_aggregate_check(nq, O, lk, la, fk, b)
        print("join / join_aggregate / group_by", nl, nr, groups, "ok", flush=True)
    # two payload columns: row numbers in the slots, streaming gather
    nl, nr = 20_000, 150_000
    L = pa.RecordBatch.from_arrays([pa.array(np.arange(nl) * 7), pa.array(rng.integers(0, 9, nl)), pa.array(rng.random(nl))], names=["k", "p", "q"])
    R = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, nl * 8, nr)), pa.array(rng.random(nr))], names=["fk", "b"])
    j = nq.HashJoin.create(nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None),
                           nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None), [("k", "fk")], "Inner").execute()[0]
    fkv = R.column(0).to_numpy()
    m = (fkv % 7 == 0) & (fkv // 7 < nl)
    assert j.num_rows == int(m.sum()) and np.array_equal(j.column(3).to_numpy(), fkv[m])
    assert np.array_equal(j.column(1).to_numpy(), L.column(1).to_numpy()[fkv[m] // 7])
    print("join, two payload columns ok", flush=True)
if "misc" in which:
    parts = [nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 9, m), mask=rng.random(m) < 0.2),
                                                                   pa.array(["x" * int(i % 4) for i in range(m)])], names=["a", "s"]))
             for m in (17, 1000, 1)]
    c = nq.DeviceTable.concat(parts)
    assert c.num_rows == 1018 and c.slice(5, 100).to_arrow().num_rows == 100
    print("concat / slice ok", flush=True)
print("sanitize_run: all ok", flush=True)
