"""`NaiveDB.run_sql` front end (sql.py + db.py), CPU side: the SQL surface SQLPlanner accepts (src/sql/planner.rs:45-380),
the logical plan shape it builds, the output names/types of `data_field` (logical_plan/expression.rs:236-331) and the
physical nodes QueryPlanner creates (src/planner/mod.rs:42-227).  No GPU call is made: plans are built, not executed."""
import os
import sys

import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import nqe_b200 as nq  # noqa: E402
from importlib import import_module  # noqa: E402

S = import_module("naive-query-engine_b200.sql")
D = import_module("naive-query-engine_b200.db")


def db():
    d = nq.NaiveDB()
    t1 = pa.schema([("id", pa.int64()), ("name", pa.utf8()), ("age", pa.int64()), ("score", pa.float64())])
    d.create_memory_table("t1", t1, [])
    d.create_memory_table("employee", pa.schema([("id", pa.int64()), ("name", pa.utf8()), ("department_id", pa.int64()), ("rank", pa.int64())]), [])
    d.create_memory_table("rank", pa.schema([("id", pa.int64()), ("rank_name", pa.utf8())]), [])
    d.create_memory_table("department", pa.schema([("id", pa.int64()), ("department_name", pa.utf8())]), [])
    return d


def chain(p):
    out = []
    while True:
        out.append(type(p).__name__)
        kids = p.children()
        if not kids:
            return out
        p = kids[0]


def test_parser_precedence_and_literals():
    s = S.parse("select id, age+100*2 from T1 where id < 9 and name = 'al''ice' or score >= 1.5e1 limit 3 offset 2")
    assert s.projection[1] == ("bin", "Plus", ("col", None, "age"), ("bin", "Multiply", ("lit", "Int64", 100), ("lit", "Int64", 2)))
    assert s.from_ == [("t1", [])]  # unquoted identifiers are lower-cased (normalize_ident)
    w = s.selection
    assert w[1] == "Or" and w[2][1] == "And" and w[2][3] == ("bin", "Eq", ("col", None, "name"), ("lit", "Utf8", "al'ice"))
    assert w[3] == ("bin", "GtEq", ("col", None, "score"), ("lit", "Float64", 15.0))
    assert s.limit == ("lit", "Int64", 3) and s.offset == ("lit", "Int64", 2)


def test_config1_plan_shape_and_names():  # BASELINE configs[0]
    p = db().plan("select id, age+100 from t1 where id < 9")
    assert chain(p) == ["ProjectionPlan", "SelectionPlan", "ScanPlan"]
    assert p.schema().names == ["id", "age + 100"] and p.schema().types == [pa.int64(), pa.int64()]
    pred = p.children()[0].expr
    assert isinstance(pred, nq.PhysicalBinaryExpr) and pred.op == "Lt" and pred.left.idx == 0


def test_readme_queries_plan_shapes():  # README.md:60-111
    d = db()
    p = d.plan("select id, name, age + 100 from t1 where id < 9 limit 3 offset 2")
    assert chain(p) == ["PhysicalLimitPlan", "PhysicalOffsetPlan", "ProjectionPlan", "SelectionPlan", "ScanPlan"]  # offset before limit
    p = d.plan("select id, name, rank_name, department_name from employee join rank on employee.rank = rank.id "
               "join department on employee.department_id = department.id")
    assert chain(p) == ["ProjectionPlan", "HashJoin", "HashJoin", "ScanPlan"]
    j2 = p.children()[0]
    assert j2.on == [("department_id", "id")] and j2.children()[0].on == [("rank", "id")]
    assert [e.idx for e in p.expr] == [0, 1, 5, 7]  # `id` is the FIRST field of that name (employee.id)
    p = d.plan("select count(id), sum(age), sum(score), avg(score), max(score), min(score) from t1 group by id % 3")
    assert chain(p) == ["ProjectionPlan", "PhysicalAggregatePlan", "ScanPlan"]
    assert len(p.schema()) == 0  # zero-field projection: the aggregate's batches pass through (projection.rs:46-48)
    agg = p.children()[0]
    assert [type(o).__name__ for o in agg.aggr_ops] == ["Count", "Sum", "Sum", "Avg", "Max", "Min"]
    assert isinstance(agg.group_expr[0], nq.PhysicalBinaryExpr) and agg.group_expr[0].op == "Modulos"


def test_comma_join_keys_come_from_where():  # planner.rs:305-380
    p = db().plan("select name, rank_name from employee, rank where rank.id = employee.rank and employee.id > 1")
    assert chain(p) == ["ProjectionPlan", "SelectionPlan", "HashJoin", "ScanPlan"]
    assert p.children()[0].children()[0].on == [("rank", "id")]  # re-ordered to (left table's column, right table's)
    with pytest.raises(nq.NqeError) as e:
        db().plan("select name from employee, rank")
    assert e.value.kind == "NotImplemented"


def test_join_on_with_extra_filter_and_outer_join_is_inner():  # planner.rs:238-280; join_type ignored at execution
    p = db().plan("select name from employee left join rank on employee.rank = rank.id")
    assert chain(p) == ["ProjectionPlan", "HashJoin", "ScanPlan"] and p.children()[0].join_type == "Left"
    p = db().plan("select name from employee join rank on employee.rank = rank.id and rank.id > 0")
    assert chain(p) == ["ProjectionPlan", "SelectionPlan", "HashJoin", "ScanPlan"]


@pytest.mark.parametrize("sql,kind", [
    ("select id as x from t1", "Panic"),                       # alias: unimplemented!()
    ("select (id + 1) from t1", "Panic"),                      # Expr::Nested: todo!()
    ("select -id from t1", "Panic"),                           # unary minus: unimplemented!()
    ("select id from t1 limit 1.5", "PlanError"),
    ("select foo(id) from t1", "PlanError"),                   # NoMatchFunction
    ("select sum(id + 1) from t1", "PlanError"),               # Aggregate Func should have a column in it
    ("select nope from t1", "PlanError"),                      # No field named
    ("select id from missing", "PlanError"),
    ("select id from t1 cross join rank", "NotImplemented"),   # CrossJoin: out of scope
    ("update t1 set id = 1", "Panic"),
])
def test_error_behaviour(sql, kind):
    with pytest.raises(nq.NqeError) as e:
        db().plan(sql)
    assert e.value.kind == kind, e.value


def test_wildcard_cast_and_abs_types():
    d = db()
    p = d.plan("select * from t1")
    assert p.schema().names == ["id", "name", "age", "score"]
    p = d.plan("select cast(id as double), abs(score) from t1")
    assert p.schema().names == ["id", "abs(score)"] and p.schema().types == [pa.float64(), pa.int64()]  # abs is typed Int64 (expression.rs:379-384)
    assert isinstance(p.expr[0], nq.PhysicalCastExpr) and isinstance(p.expr[1], nq.PhysicalUnaryExpr)
