"""`NaiveDB` -- the reference's public entry point (src/db.rs:13-47) over the GPU operators: `run_sql()` parses (sql.py),
builds the logical plan the way SQLPlanner does (src/sql/planner.rs:45-380), lowers it the way QueryPlanner does
(src/planner/mod.rs:42-227) onto the physical nodes of physical_plan.py -- whose execute() runs on the GPU through the
C ABI -- and executes the root (db.rs:36).  SURVEY.md 8f-1: the caller wiring around the hot path, plus the in-memory
table entry the reference has but does not expose (`Catalog::add_memory_table`, src/catalog.rs:38-49).

The optimizer is the reference's: an empty rule list (src/optimizer/mod.rs:12-28).  CrossJoin stays out of scope
(SURVEY.md 2, row 8): a join without keys raises NotImplemented.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import pyarrow as pa

from . import physical_plan as P
from . import sql as S
from ._ffi import NqeError

Field = Tuple[str, pa.DataType, bool]  # NaiveField without the qualifier (name lookups ignore it, planner/mod.rs:190-194)


@dataclass
class CsvConfig:
    """datasource/csv.rs:23-43."""
    has_header: bool = True
    delimiter: str = ","
    max_read_records: Optional[int] = 3
    batch_size: int = 1_000_000


# ------------------------------------------------------------------ logical plan (src/logical_plan/plan.rs:17-50)
@dataclass
class Logical:
    kind: str                    # TableScan | Projection | Filter | Join | Aggregate | Limit | Offset
    schema: List[Field]
    input: Optional["Logical"] = None
    right: Optional["Logical"] = None
    source: object = None
    exprs: Sequence = ()
    predicate: object = None
    on: Sequence = ()
    join_type: str = "Inner"
    group_expr: Sequence = ()
    aggr_expr: Sequence = ()
    n: int = 0


_OP_SYM = {"Eq": "=", "NotEq": "!=", "Lt": "<", "LtEq": "<=", "Gt": ">", "GtEq": ">=", "Plus": "+", "Minus": "-",
           "Multiply": "*", "Divide": "/", "Modulos": "%", "And": "and", "Or": "or"}
_LIT_TYPE = {"Boolean": pa.bool_(), "Int64": pa.int64(), "UInt64": pa.uint64(), "Float64": pa.float64(), "Utf8": pa.utf8(),
             "Null": pa.null()}
_SQL_TYPE = {"BOOLEAN": pa.bool_(), "SMALLINT": pa.int16(), "INT": pa.int32(), "INTEGER": pa.int32(), "BIGINT": pa.int64(),
             "FLOAT": pa.float32(), "REAL": pa.float32(), "DOUBLE": pa.float64(), "CHAR": pa.utf8(), "VARCHAR": pa.utf8(),
             "DATE": pa.date32()}


def _field_by_name(schema: List[Field], name: str) -> Field:
    for f in schema:  # field_with_unqualified_name: the first match (schema.rs:116-131)
        if f[0] == name:
            return f
    raise NqeError(6, f"No field named '{name}'")


def _lit_text(e) -> str:
    if e[2] is None:
        return "null"
    if e[1] == "Boolean":
        return "true" if e[2] else "false"
    return str(e[2])


def data_field(e, schema: List[Field]) -> Field:
    """LogicalExpr::data_field (logical_plan/expression.rs:54-92, 236-331, 365-470): output name and type."""
    k = e[0]
    if k == "col":
        return _field_by_name(schema, e[2])
    if k == "lit":
        return (_lit_text(e), _LIT_TYPE[e[1]], True)
    if k == "bin":
        left = data_field(e[2], schema)
        right = _lit_text(e[3]) if e[3][0] == "lit" else data_field(e[3], schema)[0]
        name = f"{left[0]} {_OP_SYM[e[1]]} {right}"
        boolean = e[1] in ("Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq", "And", "Or")
        return (name, pa.bool_() if boolean else left[1], True)   # arithmetic: the LEFT operand's type
    if k == "un":
        return (f"abs({data_field(e[2], schema)[0]})", pa.int64(), True)  # hard-coded Int64 (expression.rs:379-384)
    if k == "cast":
        return (data_field(e[1], schema)[0], e[2], True)
    if k == "agg":
        arg = data_field(e[2], schema)
        return (f"{e[1]}({arg[0]})", arg[1], True)
    raise NqeError(2, "Wildcard not supported in logical plan")


class SQLPlanner:
    """src/sql/planner.rs."""

    def __init__(self, catalog: Dict[str, P.MemTable]):
        self.catalog = catalog

    def statement_to_plan(self, st: S.Select) -> Logical:
        plans = [self._table_with_joins(rel, joins) for rel, joins in st.from_]
        plan = self._selection(st.selection, plans)
        items = []
        for it in st.projection:
            e = self._expr(it)
            items.extend([("col", None, f[0]) for f in plan.schema] if e[0] == "wildcard" else [e])
        aggr = [e for e in items if e[0] == "agg"]
        proj = [e for e in items if e[0] != "agg"]
        if aggr:  # plan_from_aggregate (:86-106)
            groups = [self._expr(g) for g in st.group_by]
            schema = [data_field(g, plan.schema) for g in groups] + [data_field(a, plan.schema) for a in aggr]
            plan = Logical("Aggregate", schema, input=plan, group_expr=groups, aggr_expr=aggr)
        # plan_from_projection with the NON-aggregate select items (:78, 296-303)
        plan = Logical("Projection", [data_field(e, plan.schema) for e in proj], input=plan, exprs=proj)
        if st.offset is not None:  # offset before limit (:49-52)
            plan = Logical("Offset", plan.schema, input=plan, n=self._count(st.offset, "Offset"))
        if st.limit is not None:
            plan = Logical("Limit", plan.schema, input=plan, n=self._count(st.limit, "LIMIT"))
        return plan

    def _count(self, e, what: str) -> int:
        e = self._expr(e)
        if e[0] == "lit" and e[1] == "Int64" and e[2] is not None:
            return int(e[2])
        raise NqeError(6, f"Unexpected expression for {what} clause")

    def _scan(self, name: str) -> Logical:
        if name not in self.catalog:
            raise NqeError(6, f"No table named '{name}'")  # ErrorCode::NoSuchTable
        src = self.catalog[name]
        return Logical("TableScan", [(f.name, f.type, f.nullable) for f in src.schema()], source=src)

    def _join(self, left: Logical, right: Logical, jt: str, on) -> Logical:
        if not on:  # DataFrame::join -> LogicalPlan::CrossJoin (dataframe.rs:113-121)
            raise NqeError(4, "CrossJoin is outside the hot path (SURVEY.md 2)")
        return Logical("Join", left.schema + right.schema, input=left, right=right, on=list(on), join_type=jt)

    def _table_with_joins(self, rel: str, joins) -> Logical:
        left = self._scan(rel)
        for jt, tname, on in joins:
            right = self._scan(tname)
            if on is None:
                left = self._join(left, right, jt, [])
                continue
            keys, filters = [], []
            _extract_join_keys(self._expr(on), keys, filters)
            if filters and jt != "Inner":
                raise NqeError(4, "outer join with a non-equality condition")
            left = self._join(left, right, jt, keys)
            if filters:
                pred = filters[0]
                for f in filters[1:]:
                    pred = ("bin", "And", pred, f)
                left = Logical("Filter", left.schema, input=left, predicate=pred)
        return left

    def _selection(self, where, plans: List[Logical]) -> Logical:
        if where is None:
            if len(plans) == 1:
                return plans[0]
            raise NqeError(4, "comma join without a WHERE clause")  # CROSS JOIN NOT SUPPORTED YET (:376-379)
        pred = self._expr(where)
        possible = []
        _possible_join_keys(pred, possible)
        used = set()
        left = plans[0]
        for right in plans[1:]:
            ln, rn = [f[0] for f in left.schema], [f[0] for f in right.schema]
            keys = []
            for a, b in possible:
                if a[2] in ln and b[2] in rn:
                    keys.append((a, b))
                elif b[2] in ln and a[2] in rn:
                    keys.append((b, a))
            if not keys:
                raise NqeError(4, "comma join without join keys in the WHERE clause")
            left = self._join(left, right, "Inner", keys)
            used.update(keys)
        rest = _remove_join_expressions(pred, used)
        return left if rest is None else Logical("Filter", left.schema, input=left, predicate=rest)

    def _expr(self, e):
        k = e[0]
        if k in ("col", "lit", "wildcard"):
            return e
        if k == "bin":
            return ("bin", e[1], self._expr(e[2]), self._expr(e[3]))
        if k == "un":
            return ("un", e[1], self._expr(e[2]))
        if k == "cast":
            if e[2] not in _SQL_TYPE:
                if e[2] in ("TIMESTAMP", "DECIMAL"):
                    raise NqeError(5, "not yet implemented")  # todo!() (:585-587)
                raise NqeError(6, f"Unsupported SQL type {e[2]}")
            return ("cast", self._expr(e[1]), _SQL_TYPE[e[2]])
        if k == "call":
            args = [self._expr(a) for a in e[2]]
            name = e[1]
            if name == "abs":
                if len(args) != 1:
                    raise NqeError(6, "Scalar Func only has one parameter")
                return ("un", "Abs", args[0])
            if name in ("count", "sum", "avg", "min", "max"):
                if len(args) != 1:
                    raise NqeError(6, "Aggregate Func Now only Support One parameter")
                return ("agg", name, args[0])
            raise NqeError(6, f"Not find match func: {name}")  # ErrorCode::NoMatchFunction
        raise NqeError(5, "not yet implemented")


def _extract_join_keys(e, keys, filters):
    """extract_join_keys (planner.rs:606-640)"""
    if e[0] == "bin":
        if e[1] == "Eq":
            if e[2][0] == "col" and e[3][0] == "col":
                keys.append((e[2], e[3]))
            else:
                filters.append(e)
        elif e[1] == "And":
            _extract_join_keys(e[2], keys, filters)
            _extract_join_keys(e[3], keys, filters)
        elif e[2][0] == "col" or e[3][0] == "col":
            filters.append(e)
        else:
            _extract_join_keys(e[2], keys, filters)
            _extract_join_keys(e[3], keys, filters)
    else:
        filters.append(e)


def _possible_join_keys(e, out):
    if e[0] == "bin" and e[1] == "Eq" and e[2][0] == "col" and e[3][0] == "col":
        out.append((e[2], e[3]))
    elif e[0] == "bin" and e[1] == "And":
        _possible_join_keys(e[2], out)
        _possible_join_keys(e[3], out)


def _remove_join_expressions(e, used):
    if e[0] == "bin" and e[1] == "Eq" and e[2][0] == "col" and e[3][0] == "col":
        return None if (e[2], e[3]) in used or (e[3], e[2]) in used else e
    if e[0] == "bin" and e[1] == "And":
        l, r = _remove_join_expressions(e[2], used), _remove_join_expressions(e[3], used)
        if l is not None and r is not None:
            return ("bin", "And", l, r)
        return l if l is not None else r
    return e


class QueryPlanner:
    """src/planner/mod.rs:42-227: LogicalPlan -> PhysicalPlan, LogicalExpr -> PhysicalExpr."""

    @staticmethod
    def create_physical_plan(plan: Logical) -> P.PhysicalPlan:
        cpp = QueryPlanner.create_physical_plan
        cpe = QueryPlanner.create_physical_expression
        if plan.kind == "TableScan":
            return P.ScanPlan.create(plan.source, None)
        if plan.kind == "Projection":
            inp = cpp(plan.input)
            exprs = [cpe(e, plan.input.schema) for e in plan.exprs]
            schema = pa.schema([pa.field(n, t, nullable) for n, t, nullable in plan.schema])
            return P.ProjectionPlan.create(inp, schema, exprs)
        if plan.kind == "Limit":
            return P.PhysicalLimitPlan.create(cpp(plan.input), plan.n)
        if plan.kind == "Offset":
            return P.PhysicalOffsetPlan.create(cpp(plan.input), plan.n)
        if plan.kind == "Join":
            schema = pa.schema([pa.field(n, t, nullable) for n, t, nullable in plan.schema])
            on = [(a[2], b[2]) for a, b in plan.on]
            return P.HashJoin.create(cpp(plan.input), cpp(plan.right), on, plan.join_type, schema)
        if plan.kind == "Filter":
            pred = cpe(plan.predicate, plan.schema)  # `create_physical_expression(&filter.predicate, plan)` (:91)
            return P.SelectionPlan.create(cpp(plan.input), pred)
        if plan.kind == "Aggregate":
            groups = [cpe(g, plan.input.schema) for g in plan.group_expr]
            ops = []
            cls = {"count": P.Count, "sum": P.Sum, "avg": P.Avg, "min": P.Min, "max": P.Max}
            for a in plan.aggr_expr:
                arg = cpe(a[2], plan.input.schema)
                if not isinstance(arg, P.ColumnExpr):
                    raise NqeError(6, "Aggregate Func should have a column in it")
                ops.append(cls[a[1]].create(arg))
            return P.PhysicalAggregatePlan.create(groups, ops, cpp(plan.input))
        raise NqeError(5, f"not implemented: {plan.kind}")

    @staticmethod
    def create_physical_expression(e, schema: List[Field]) -> P.PhysicalExpr:
        cpe = QueryPlanner.create_physical_expression
        k = e[0]
        if k == "col":
            for idx, f in enumerate(schema):
                if f[0] == e[2]:
                    return P.ColumnExpr.try_create(None, idx)
            raise NqeError(7, f"column `{e[2]}` not exists")  # ErrorCode::ColumnNotExists
        if k == "lit":
            sv = P.ScalarValue
            mk = {"Boolean": sv.Boolean, "Int64": sv.Int64, "UInt64": sv.UInt64, "Float64": sv.Float64, "Utf8": sv.Utf8}
            if e[1] == "Null":
                return P.PhysicalLiteralExpr.create(P.ScalarValue("Null", None))
            return P.PhysicalLiteralExpr.create(mk[e[1]](e[2]))
        if k == "bin":
            return P.PhysicalBinaryExpr.create(cpe(e[2], schema), e[1], cpe(e[3], schema))
        if k == "un":
            return P.PhysicalUnaryExpr.create(cpe(e[2], schema), e[1], "todo", pa.int32())
        if k == "cast":
            return P.PhysicalCastExpr.create(cpe(e[1], schema), e[2])
        raise NqeError(5, "not yet implemented")  # Alias / Not / AggregateFunction / Wildcard: todo!()


class NaiveDB:
    """src/db.rs:13-47."""

    def __init__(self):
        self.catalog: Dict[str, P.MemTable] = {}

    def create_csv_table(self, table: str, csv_file: str, csv_conf: Optional[CsvConfig] = None) -> None:
        conf = csv_conf or CsvConfig()
        self.catalog[table] = P.CsvTable.try_create(csv_file, conf.has_header, conf.delimiter)

    def create_memory_table(self, table: str, schema: pa.Schema, batches: Sequence[pa.RecordBatch]) -> None:
        """Catalog::add_memory_table (src/catalog.rs:38-49): how batches larger than the CSV reader's 1e6-row cap enter."""
        self.catalog[table] = P.MemTable.try_create(schema, list(batches))

    def plan(self, sql: str) -> P.PhysicalPlan:
        logical = SQLPlanner(self.catalog).statement_to_plan(S.parse(sql))
        # Optimizer::default().optimize: no rules (src/optimizer/mod.rs:22-28)
        return QueryPlanner.create_physical_plan(logical)

    def run_sql(self, sql: str) -> List[pa.RecordBatch]:
        return self.plan(sql).execute()
