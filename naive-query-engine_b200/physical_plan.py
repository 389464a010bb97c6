"""Host-side mirror of the reference's physical_plan interface, driving the
CUDA kernels through the C ABI.

Same names, argument meaning and error behaviour as the reference:
    trait PhysicalPlan {schema, execute, children}      src/physical_plan/plan.rs:14-23
    trait PhysicalExpr {evaluate}                        src/physical_plan/expression/mod.rs:25-31
    trait AggregateOperator                              src/physical_plan/aggregate/mod.rs:225-235
    ScanPlan / SelectionPlan / ProjectionPlan / HashJoin / PhysicalAggregatePlan /
    PhysicalLimitPlan / PhysicalOffsetPlan               src/physical_plan/*.rs

`execute()` returns host Arrow batches like the reference.  Between operators
the data stays in HBM (`execute_device()`), and two adjacent-node patterns are
fused into one kernel pass:
    ProjectionPlan(SelectionPlan(x))            -> nqe_filter_project
    PhysicalAggregatePlan(HashJoin(l, r))       -> nqe_join_aggregate (bare-column group key)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import pyarrow as pa

from . import _ffi
from ._ffi import Agg, Expr, ExprNode, NqeError
from .device import Context, DeviceTable, nqe_dtype

# Operator, src/logical_plan/expression.rs:335-362
OPERATORS = ["Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq", "Plus", "Minus", "Multiply", "Divide", "Modulos", "And", "Or"]
# UnaryOperator, src/logical_plan/expression.rs:392-422 (implemented subset, unary.rs:92-96)
UNARY = ["Abs", "Sin", "Cos", "Tan"]
_TODO_UNARY = ["Trim", "LTrim", "RTrim", "CharacterLength", "Lower", "Upper", "Repeat", "Replace", "Reverse", "Substr"]


# --------------------------------------------------------------------------
# ScalarValue, src/logical_plan/expression.rs:174-187
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class ScalarValue:
    kind: str  # "Null" | "Boolean" | "Float64" | "Int64" | "UInt64" | "Utf8"
    value: object = None

    @staticmethod
    def Null():
        return ScalarValue("Null")

    @staticmethod
    def Boolean(v):
        return ScalarValue("Boolean", v)

    @staticmethod
    def Float64(v):
        return ScalarValue("Float64", v)

    @staticmethod
    def Int64(v):
        return ScalarValue("Int64", v)

    @staticmethod
    def UInt64(v):
        return ScalarValue("UInt64", v)

    @staticmethod
    def Utf8(v):
        return ScalarValue("Utf8", v)


_SCALAR_DTYPE = {"Boolean": _ffi.BOOL, "Float64": _ffi.FLOAT64, "Int64": _ffi.INT64, "UInt64": _ffi.UINT64,
                 "Utf8": _ffi.UTF8, "Null": 0}


# --------------------------------------------------------------------------
# PhysicalExpr tree -> postfix nqe_expr_node program
# --------------------------------------------------------------------------
class PhysicalExpr:
    def lower(self, names: Sequence[str], out: List[ExprNode]):
        raise NotImplementedError

    def to_expr(self, names: Sequence[str]) -> Tuple[Expr, object]:
        """Postfix lowering for the C ABI.  Expression trees are immutable after create(), so the lowered form
        is kept per input schema (a plan node is usually executed against the same schema every time)."""
        key = tuple(names)
        cached = getattr(self, "_lowered", None)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        nodes: List[ExprNode] = []
        self.lower(names, nodes)
        arr = (ExprNode * len(nodes))(*nodes)
        ex = Expr(arr, len(nodes), 0)
        self._lowered = (key, ex, arr)
        return ex, arr

    def evaluate(self, batch) -> pa.Array:
        """PhysicalExpr::evaluate(&RecordBatch).into_array() -- runs on the GPU."""
        t = batch if isinstance(batch, DeviceTable) else DeviceTable.from_arrow(batch)
        out = _filter_project(t, None, [self], ["expr"])
        return out.to_arrow().column(0)


class ColumnExpr(PhysicalExpr):
    """column.rs:18-57: by idx if set, else first field whose name matches."""

    def __init__(self, name: Optional[str], idx: Optional[int]):
        self.name, self.idx = name, idx

    @staticmethod
    def try_create(name: Optional[str], idx: Optional[int]) -> "ColumnExpr":
        if name is None and idx is None:
            raise NqeError(7, "ColumnExpr must has name or idx")
        return ColumnExpr(name, idx)

    def resolve(self, names: Sequence[str]) -> int:
        if self.idx is not None:
            return self.idx
        for i, n in enumerate(names):
            if n == self.name:
                return i
        raise NqeError(7, "ColumnExpr must has name or idx")

    def lower(self, names, out):
        out.append(ExprNode(0, 0, self.resolve(names), 0, 0, 0))

    def __repr__(self):
        return f"ColumnExpr(name={self.name!r}, idx={self.idx!r})"


class PhysicalLiteralExpr(PhysicalExpr):
    """literal.rs:17-35: ColumnValue::Const(v, n) -- kept as an immediate on the GPU."""

    def __init__(self, literal: ScalarValue):
        self.literal = literal

    @staticmethod
    def create(literal: ScalarValue) -> "PhysicalLiteralExpr":
        return PhysicalLiteralExpr(literal)

    def lower(self, names, out):
        lit = self.literal
        n = ExprNode(1, 0, 0, _SCALAR_DTYPE[lit.kind], 1 if lit.value is None else 0, 0)
        if lit.kind == "Utf8" and lit.value is not None:
            # the bytes must outlive the call: they hang off this expression object (nqe.h: value.u64 = address, reserved = length)
            self._utf8 = C.create_string_buffer(str(lit.value).encode("utf-8"))
            n.reserved = len(self._utf8.raw) - 1
            n.value.u64 = C.addressof(self._utf8)
        elif lit.value is not None:
            if lit.kind == "Float64":
                n.value.f64 = float(lit.value)
            elif lit.kind == "UInt64":
                n.value.u64 = int(lit.value)
            elif lit.kind == "Boolean":
                n.value.u64 = 1 if lit.value else 0
            elif lit.kind == "Int64":
                n.value.i64 = int(lit.value)
        out.append(n)

    def __repr__(self):
        return f"PhysicalLiteralExpr({self.literal})"


class PhysicalBinaryExpr(PhysicalExpr):
    """binary.rs:91-155."""

    def __init__(self, left, op: str, right):
        if op not in OPERATORS:
            raise ValueError(op)
        self.left, self.op, self.right = left, op, right

    @staticmethod
    def create(left, op: str, right) -> "PhysicalBinaryExpr":
        return PhysicalBinaryExpr(left, op, right)

    def lower(self, names, out):
        self.left.lower(names, out)
        self.right.lower(names, out)
        out.append(ExprNode(2, OPERATORS.index(self.op), 0, 0, 0, 0))

    def __repr__(self):
        return f"PhysicalBinaryExpr({self.left!r}, {self.op}, {self.right!r})"


class PhysicalUnaryExpr(PhysicalExpr):
    """unary.rs:46-108 (Tan evaluates cos, :96; string functions are todo!())."""

    def __init__(self, expr, func: str, name: str, return_type=None):
        self.expr, self.func, self.name, self.return_type = expr, func, name, return_type

    @staticmethod
    def create(expr, func: str, name: str, return_type=None) -> "PhysicalUnaryExpr":
        return PhysicalUnaryExpr(expr, func, name, return_type)

    def lower(self, names, out):
        if self.func in _TODO_UNARY:
            raise NqeError(5, "not yet implemented")  # todo!() in the reference
        self.expr.lower(names, out)
        out.append(ExprNode(3, UNARY.index(self.func), 0, 0, 0, 0))


class PhysicalCastExpr(PhysicalExpr):
    """cast.rs:45-87: every arm is todo!() in the reference -- evaluating a CAST panics."""

    def __init__(self, expr, data_type):
        self.expr, self.data_type = expr, data_type

    @staticmethod
    def create(expr, data_type) -> "PhysicalCastExpr":
        return PhysicalCastExpr(expr, data_type)

    def lower(self, names, out):
        raise NqeError(5, "not yet implemented")


# --------------------------------------------------------------------------
# AggregateOperator impls, aggregate/{count,sum,avg,max,min}.rs
# --------------------------------------------------------------------------
class AggregateOperator:
    OP = -1
    FN = ""

    def __init__(self, col_expr: ColumnExpr):
        if not isinstance(col_expr, ColumnExpr):
            # planner/mod.rs:104-163: the argument must downcast to ColumnExpr
            raise NqeError(6, "Aggregate Func should have a column in it")
        self.col_expr = col_expr

    @classmethod
    def create(cls, col_expr: ColumnExpr):
        return cls(col_expr)

    def data_field(self, names: Sequence[str]) -> pa.Field:
        """e.g. sum.rs:57-82: `sum(col)`, Float64 (count: UInt64), non-nullable."""
        idx = self.col_expr.resolve(names)
        return pa.field(f"{self.FN}({names[idx]})", pa.uint64() if self.OP == 0 else pa.float64(), nullable=False)


class Count(AggregateOperator):
    OP, FN = 0, "count"


class Sum(AggregateOperator):
    OP, FN = 1, "sum"


class Avg(AggregateOperator):
    OP, FN = 2, "avg"


class Min(AggregateOperator):
    OP, FN = 3, "min"


class Max(AggregateOperator):
    OP, FN = 4, "max"


# --------------------------------------------------------------------------
# helpers over the C ABI
# --------------------------------------------------------------------------
def _filter_project(t: DeviceTable, predicate: Optional[PhysicalExpr], exprs: Sequence[PhysicalExpr],
                    out_names: Sequence[str]) -> DeviceTable:
    ctx = t.ctx
    keep = []
    pred_ptr = None
    if predicate is not None:
        pe, arr = predicate.to_expr(t.names)
        keep.append(arr)
        pred_ptr = C.pointer(pe)
    n = len(exprs)
    earr = (Expr * max(n, 1))()
    for i, e in enumerate(exprs):
        ex, arr = e.to_expr(t.names)
        keep.append(arr)
        earr[i] = ex
    h = C.c_void_p()
    ctx.check(ctx.lib.nqe_filter_project(ctx.h, t.h, pred_ptr, earr if n else None, n, C.byref(h)))
    return DeviceTable(ctx, h, list(out_names) if n else list(t.names))


def filter_project_host(ctx, names: Sequence[str], dtypes: Sequence[int], host_ptrs: Sequence[int], n_rows: int,
                        predicate: Optional[PhysicalExpr], exprs: Sequence[PhysicalExpr], out_ptrs: Sequence[int],
                        out_capacity_rows: int) -> Tuple[int, List[int]]:
    """ProjectionPlan(SelectionPlan(scan)) over HOST column buffers with the result written into HOST buffers
    (nqe_filter_project_host: chunked H2D | kernel | D2H pipeline, only referenced columns are uploaded).
    host_ptrs / out_ptrs are addresses of 8-byte-value buffers (pinned memory is copied asynchronously).
    Returns (result rows, result dtypes)."""
    from ._ffi import ColumnDesc
    keep = []
    pred_ptr = None
    if predicate is not None:
        pe, arr = predicate.to_expr(list(names))
        keep.append(arr)
        pred_ptr = C.pointer(pe)
    earr = (Expr * len(exprs))()
    for i, e in enumerate(exprs):
        ex, arr = e.to_expr(list(names))
        keep.append(arr)
        earr[i] = ex
    cols = (ColumnDesc * len(names))()
    for i, (dt, ptr) in enumerate(zip(dtypes, host_ptrs)):
        cols[i].dtype, cols[i].length, cols[i].null_count, cols[i].values = dt, n_rows, 0, ptr
    outs = (ColumnDesc * len(exprs))()
    for i, ptr in enumerate(out_ptrs):
        outs[i].length, outs[i].values = out_capacity_rows, ptr
    rows = C.c_int64(0)
    ctx.check(ctx.lib.nqe_filter_project_host(ctx.h, cols, len(names), pred_ptr, earr, len(exprs), outs, C.byref(rows)))
    return int(rows.value), [int(outs[i].dtype) for i in range(len(exprs))]


# --------------------------------------------------------------------------
# PhysicalPlan nodes
# --------------------------------------------------------------------------
class PhysicalPlan:
    def schema(self) -> pa.Schema:
        raise NotImplementedError

    def children(self) -> List["PhysicalPlan"]:
        raise NotImplementedError

    def execute_device(self) -> DeviceTable:
        raise NotImplementedError

    def execute(self) -> List[pa.RecordBatch]:
        """PhysicalPlan::execute: fully materialised host Arrow batches."""
        return [self.execute_device().to_arrow()]


class MemTable:
    """datasource/memory.rs:17-57: in-memory batches; honours column projection."""

    def __init__(self, schema: pa.Schema, batches: Sequence[pa.RecordBatch]):
        self._schema, self.batches = schema, list(batches)
        self._device: Optional[DeviceTable] = None

    @staticmethod
    def try_create(schema: pa.Schema, batches: Sequence[pa.RecordBatch]) -> "MemTable":
        return MemTable(schema, batches)

    def schema(self) -> pa.Schema:
        return self._schema

    def source_name(self) -> str:
        return "MemTable"

    def scan(self, projection: Optional[Sequence[int]] = None) -> List[pa.RecordBatch]:
        if projection is None:
            return list(self.batches)
        return [pa.RecordBatch.from_arrays([b.column(i) for i in projection],
                                           names=[b.schema.names[i] for i in projection]) for b in self.batches]

    def device_table(self, ctx: Optional[Context] = None) -> DeviceTable:
        """Upload once and keep the table resident in HBM (the GPU analogue of
        the Arc-cloned batches in memory.rs:31-41)."""
        if self._device is None:
            if len(self.batches) <= 1:
                src = self.batches[0] if self.batches else pa.Table.from_batches([], schema=self._schema)
                self._device = DeviceTable.from_arrow(src, ctx)
            else:
                # every operator of the reference that sees several batches concatenates them first (concat_batches,
                # hash_join.rs:131-132,258-273; aggregate/mod.rs:143-144) or treats them one by one with the same
                # result as on the concatenation: each batch is uploaded as it is and concatenated on the device
                parts = [DeviceTable.from_arrow(b, ctx) for b in self.batches]
                self._device = DeviceTable.concat(parts)
                for t in parts:
                    t.free()
        return self._device


class CsvTable(MemTable):
    """datasource/csv.rs:46-96 via pyarrow.csv (I/O is outside the hot path)."""

    @staticmethod
    def try_create(path: str, has_header: bool = True, delimiter: str = ",") -> "CsvTable":
        import pyarrow.csv as pcsv
        tbl = pcsv.read_csv(path, parse_options=pcsv.ParseOptions(delimiter=delimiter),
                            read_options=pcsv.ReadOptions(autogenerate_column_names=not has_header))
        batch = tbl.combine_chunks().to_batches()[0] if tbl.num_rows else pa.RecordBatch.from_pylist([], schema=tbl.schema)
        return CsvTable(batch.schema, [batch])

    def source_name(self) -> str:
        return "CsvTable"


class ScanPlan(PhysicalPlan):
    """scan.rs:17-49."""

    def __init__(self, source: MemTable, projection: Optional[Sequence[int]]):
        self.source, self.projection = source, projection

    @staticmethod
    def create(source: MemTable, projection: Optional[Sequence[int]] = None) -> "ScanPlan":
        return ScanPlan(source, projection)

    def schema(self) -> pa.Schema:
        return self.source.schema()

    def children(self):
        return []

    def execute_device(self) -> DeviceTable:
        t = self.source.device_table()
        if self.projection is None:
            return t
        exprs = [ColumnExpr(None, i) for i in self.projection]
        return _filter_project(t, None, exprs, [t.names[i] for i in self.projection])

    def execute(self):
        return self.source.scan(self.projection)


class SelectionPlan(PhysicalPlan):
    """selection.rs:22-112."""

    def __init__(self, input: PhysicalPlan, expr: PhysicalExpr):
        self.input, self.expr = input, expr

    @staticmethod
    def create(input: PhysicalPlan, expr: PhysicalExpr) -> "SelectionPlan":
        return SelectionPlan(input, expr)

    def schema(self):
        return self.input.schema()

    def children(self):
        return [self.input]

    def execute_device(self) -> DeviceTable:
        t = self.input.execute_device()
        return _filter_project(t, self.expr, [], [])


class ProjectionPlan(PhysicalPlan):
    """projection.rs:17-74.  A zero-field schema passes the input through
    (how aggregate results flow out, projection.rs:46-48)."""

    def __init__(self, input: PhysicalPlan, schema: pa.Schema, expr: Sequence[PhysicalExpr]):
        self.input, self._schema, self.expr = input, schema, list(expr)

    @staticmethod
    def create(input: PhysicalPlan, schema: pa.Schema, expr: Sequence[PhysicalExpr]) -> "ProjectionPlan":
        return ProjectionPlan(input, schema, expr)

    def schema(self):
        return self._schema

    def children(self):
        return [self.input]

    def execute_device(self) -> DeviceTable:
        if len(self._schema) == 0:
            return self.input.execute_device()
        names = list(self._schema.names)
        if isinstance(self.input, SelectionPlan):  # fused filter -> project, one pass over HBM
            t = self.input.input.execute_device()
            out = _filter_project(t, self.input.expr, self.expr, names)
        else:
            t = self.input.execute_device()
            out = _filter_project(t, None, self.expr, names)
        # RecordBatch::try_new(schema, columns).unwrap() (projection.rs:65): dtype mismatch panics
        for f, dt in zip(self._schema, out.dtypes()):
            if nqe_dtype(f.type) != dt:
                raise NqeError(5, f"column types must match schema types, expected {f.type} for `{f.name}`")
        return out


class HashJoin(PhysicalPlan):
    """hash_join.rs:43-285.  `on`: list of (left Column name, right Column name);
    only on[0] is used, join_type is ignored (always INNER), as in the reference."""

    def __init__(self, left, right, on: Sequence[Tuple[str, str]], join_type: str, schema: pa.Schema):
        self.left, self.right, self.on, self.join_type, self._schema = left, right, list(on), join_type, schema

    @staticmethod
    def create(left, right, on, join_type="Inner", schema: Optional[pa.Schema] = None) -> "HashJoin":
        if schema is None:
            schema = pa.schema(list(left.schema()) + list(right.schema()))
        return HashJoin(left, right, on, join_type, schema)

    def schema(self):
        return self._schema

    def children(self):
        return [self.left, self.right]

    def _keys(self, lt: DeviceTable, rt: DeviceTable) -> Tuple[int, int]:
        if not self.on:
            raise NqeError(6, "Inner Join on Conditions can't not be empty")
        lname, rname = self.on[0]
        lk = ColumnExpr(lname, None).resolve(lt.names)
        rk = ColumnExpr(rname, None).resolve(rt.names)
        return lk, rk

    def execute_device(self) -> DeviceTable:
        if not self.on:
            raise NqeError(6, "Inner Join on Conditions can't not be empty")
        lt = self.left.execute_device()
        rt = self.right.execute_device()
        lk, rk = self._keys(lt, rt)
        ctx = lt.ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_join(ctx.h, lt.h, rt.h, lk, rk, C.byref(h)))
        return DeviceTable(ctx, h, list(self._schema.names))


class PhysicalAggregatePlan(PhysicalPlan):
    """aggregate/mod.rs:29-222."""

    def __init__(self, group_expr: Sequence[PhysicalExpr], aggr_ops: Sequence[AggregateOperator], input: PhysicalPlan):
        self.group_expr, self.aggr_ops, self.input = list(group_expr), list(aggr_ops), input

    @staticmethod
    def create(group_expr, aggr_ops, input) -> "PhysicalAggregatePlan":
        return PhysicalAggregatePlan(group_expr, aggr_ops, input)

    def schema(self):
        return self.input.schema()  # aggregate/mod.rs:44,105-107: the INPUT's schema

    def children(self):
        return [self.input]

    def _aggs(self, names):
        n = len(self.aggr_ops)
        arr = (Agg * max(n, 1))()
        for i, op in enumerate(self.aggr_ops):
            arr[i] = Agg(op.OP, op.col_expr.resolve(names))
        return arr, n

    def execute_device(self) -> DeviceTable:
        inp = self.input
        # fused join -> aggregate when the (single used) group expr is a bare column
        if isinstance(inp, HashJoin) and self.group_expr and isinstance(self.group_expr[0], ColumnExpr) and inp.on:
            lt = inp.left.execute_device()
            rt = inp.right.execute_device()
            lk, rk = inp._keys(lt, rt)
            names = list(inp.schema().names)
            out_names = [op.data_field(names).name for op in self.aggr_ops]
            aggs, n = self._aggs(names)
            g = self.group_expr[0].resolve(names)
            ctx = lt.ctx
            UTF8 = 5
            jdt = list(lt.dtypes()) + list(rt.dtypes())
            if lt.dtypes()[lk] != UTF8 and jdt[g] != UTF8:  # Utf8 keys go through the dictionary path, unfused
                h = C.c_void_p()
                ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, lk, rk, g, aggs, n, C.byref(h)))
                return DeviceTable(ctx, h, out_names)
        t = inp.execute_device()
        names = t.names
        out_names = [op.data_field(names).name for op in self.aggr_ops]
        aggs, n = self._aggs(names)
        ctx = t.ctx
        keep = None
        gptr = None
        if self.group_expr:  # only group_expr[0] is used (aggregate/mod.rs:141-146)
            ge, keep = self.group_expr[0].to_expr(names)
            gptr = C.pointer(ge)
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, t.h, gptr, aggs, n, C.byref(h)))
        return DeviceTable(ctx, h, out_names)


class PhysicalLimitPlan(PhysicalPlan):
    """limit.rs:17-66: first n rows."""

    def __init__(self, input, n: int):
        self.input, self.n = input, n

    @staticmethod
    def create(input, n: int) -> "PhysicalLimitPlan":
        return PhysicalLimitPlan(input, n)

    def schema(self):
        return self.input.schema()

    def children(self):
        return [self.input]

    def execute_device(self):
        t = self.input.execute_device()
        return t.slice(0, min(self.n, t.num_rows))


class PhysicalOffsetPlan(PhysicalPlan):
    """offset.rs:17-68: skip n rows."""

    def __init__(self, input, n: int):
        self.input, self.n = input, n

    @staticmethod
    def create(input, n: int) -> "PhysicalOffsetPlan":
        return PhysicalOffsetPlan(input, n)

    def schema(self):
        return self.input.schema()

    def children(self):
        return [self.input]

    def execute_device(self):
        t = self.input.execute_device()
        off = min(self.n, t.num_rows)
        return t.slice(off, t.num_rows - off)
