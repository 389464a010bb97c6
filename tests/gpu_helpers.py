"""Bridges between the oracle's data model and the product's host API."""
import numpy as np
import pyarrow as pa

import nqe_b200 as nq
from oracle import oracle as O

_PA = {"bool": pa.bool_(), "i64": pa.int64(), "u64": pa.uint64(), "f64": pa.float64(), "utf8": pa.utf8()}
_SV = {"bool": nq.ScalarValue.Boolean, "i64": nq.ScalarValue.Int64, "u64": nq.ScalarValue.UInt64,
       "f64": nq.ScalarValue.Float64, "utf8": nq.ScalarValue.Utf8}
_UN = {"abs": "Abs", "sin": "Sin", "cos": "Cos", "tan": "Tan"}
_AGG = {"count": nq.Count, "sum": nq.Sum, "avg": nq.Avg, "min": nq.Min, "max": nq.Max}


def to_arrow(b: O.Batch) -> pa.RecordBatch:
    arrays = []
    for c in b.cols:
        mask = None if c.valid is None else (c.valid == 0)
        if c.dtype == "bool":
            arrays.append(pa.array(c.values.astype(bool), type=pa.bool_(), mask=mask))
        elif c.dtype == "utf8":
            arrays.append(pa.array(list(c.values), type=pa.utf8(), mask=mask))
        else:
            arrays.append(pa.array(c.values, type=_PA[c.dtype], mask=mask))
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema([pa.field(n, a.type) for n, a in zip(b.names, arrays)]))


def from_arrow(rb: pa.RecordBatch) -> O.Batch:
    cols = []
    inv = {v: k for k, v in _PA.items()}
    for i in range(rb.num_columns):
        a = rb.column(i)
        cols.append(O.col(inv[a.type], a.to_pylist()))
    return O.Batch(list(rb.schema.names), cols)


def expr(e) -> nq.PhysicalExpr:
    k = e[0]
    if k == "col":
        return nq.ColumnExpr.try_create(None, e[1])
    if k == "lit":
        return nq.PhysicalLiteralExpr.create(_SV[e[1]](e[2]))
    if k == "bin":
        return nq.PhysicalBinaryExpr.create(expr(e[2]), e[1], expr(e[3]))
    if k == "un":
        return nq.PhysicalUnaryExpr.create(expr(e[2]), _UN[e[1]], e[1], pa.float64())
    raise ValueError(e)


def scan(b: O.Batch) -> nq.ScanPlan:
    rb = to_arrow(b)
    return nq.ScanPlan.create(nq.MemTable.try_create(rb.schema, [rb]), None)


def gpu_projection(b: O.Batch, exprs, pred=None, names=None) -> O.Batch:
    """[SelectionPlan +] ProjectionPlan over a ScanPlan on the GPU (one fused
    nqe_filter_project call) -> oracle Batch."""
    from importlib import import_module
    pp = import_module("naive-query-engine_b200.physical_plan")
    names = names or [O.expr_name(e, b.names) for e in exprs]
    src = scan(b).execute_device()
    t = pp._filter_project(src, expr(pred) if pred is not None else None, [expr(e) for e in exprs], names)
    return from_arrow(t.to_arrow())


def gpu_selection(b: O.Batch, pred) -> O.Batch:
    return from_arrow(nq.SelectionPlan.create(scan(b), expr(pred)).execute()[0])


def gpu_join(l: O.Batch, r: O.Batch, lkey: str, rkey: str) -> O.Batch:
    return from_arrow(nq.HashJoin.create(scan(l), scan(r), [(lkey, rkey)], "Inner").execute()[0])


def gpu_aggregate(b: O.Batch, group_expr, aggs) -> O.Batch:
    ops = [_AGG[op].create(nq.ColumnExpr.try_create(None, ci)) for op, ci in aggs]
    plan = nq.PhysicalAggregatePlan.create([expr(group_expr)] if group_expr is not None else [], ops, scan(b))
    return from_arrow(plan.execute()[0])


def gpu_join_aggregate(l: O.Batch, r: O.Batch, lkey: str, rkey: str, group_col: int, aggs) -> O.Batch:
    ops = [_AGG[op].create(nq.ColumnExpr.try_create(None, ci)) for op, ci in aggs]
    join = nq.HashJoin.create(scan(l), scan(r), [(lkey, rkey)], "Inner")
    plan = nq.PhysicalAggregatePlan.create([nq.ColumnExpr.try_create(None, group_col)], ops, join)
    return from_arrow(plan.execute()[0])
