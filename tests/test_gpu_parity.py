"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs, and against the committed golden fixtures.

Bar: bit-exact for integer / boolean / index columns, for f64 arithmetic
(+,-,*,/ are IEEE-exact) and for min/max; sum/avg within 1e-9 relative
(north_star); sin/cos within 4 ulp of glibc (CUDA libdevice is not correctly
rounded -- the reference's sin vector is bit-exact against glibc only)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests import gpu_helpers as G
from tests.helpers import assert_rows, fixtures, golden_table

pytestmark = pytest.mark.gpu
FX = fixtures()["cases"]
SUM_REL = 1e-9


def lit(v):
    return ("lit", "i64", v)


def numeric_only(b: O.Batch) -> O.Batch:
    keep = [i for i, c in enumerate(b.cols) if c.dtype != "utf8"]
    return O.Batch([b.names[i] for i in keep], [b.cols[i] for i in keep])


def rand_col(rng, dtype, n, null_frac=0.0, lo=-50, hi=50):
    if dtype == "i64":
        v = rng.integers(lo, hi, n).astype(np.int64)
    elif dtype == "u64":
        v = rng.integers(0, hi, n).astype(np.uint64)
    elif dtype == "f64":
        v = np.round(rng.normal(0, 10, n), 3)
    else:
        v = rng.integers(0, 2, n).astype(np.uint8)
    valid = None
    if null_frac > 0:
        valid = (rng.random(n) >= null_frac).astype(np.uint8)
    return O.Col(dtype, v, valid)


def same(a: O.Batch, b: O.Batch, rel=0.0, ordered=True, sort_cols=None):
    assert [c.dtype for c in a.cols] == [c.dtype for c in b.cols]
    assert_rows([list(r) for r in a.rows()], [list(r) for r in b.rows()], rel=rel, ordered=ordered, sort_cols=sort_cols)


# ---------------------------------------------------------------- golden fixtures
def test_golden_projection_vector():  # projection.rs:88-121 (numeric columns)
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_projection(t1, [("bin", "Plus", ("col", 0), lit(1))])
    assert out.cols[0].to_pylist() == FX["test_projection"]["id_plus_1"]


def test_golden_selection_vector():  # selection.rs:126-178
    t1 = numeric_only(golden_table("t1"))
    pred = ("bin", "Gt", ("bin", "Plus", ("col", 0), lit(1)), lit(5))
    out = G.gpu_selection(t1, pred)
    assert out.cols[0].to_pylist() == FX["test_selection"]["id"]


def test_golden_sql_where():  # sql/planner.rs:664-680
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_projection(t1, [("col", 0), ("col", 1)], pred=("bin", "Gt", ("col", 0), lit(1)))
    assert out.cols[0].to_pylist() == FX["sql_where_id_gt_1"]["id"]
    assert out.cols[1].to_pylist() == FX["sql_where_id_gt_1"]["age"]


def test_golden_config1():  # BASELINE.json configs[0]
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_projection(t1, [("col", 0), ("bin", "Plus", ("col", 1), lit(100))],
                           pred=("bin", "Lt", ("col", 0), lit(9)), names=FX["config1"]["names"])
    assert [list(r) for r in out.rows()] == FX["config1"]["rows"]


def test_golden_abs_sin():  # unary.rs:123-170
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_projection(t1, [("un", "abs", ("col", 2)), ("un", "sin", ("col", 2))])
    assert out.cols[0].to_pylist() == FX["test_abs_expression"]["score"]
    for got, want in zip(out.cols[1].to_pylist(), FX["test_sin_expression"]["score"]):
        assert abs(got - want) <= 4 * math.ulp(want)


def test_cos_and_tan_is_cos():  # unary.rs:92-96: `Tan` evaluates cos
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_projection(t1, [("un", "cos", ("col", 2)), ("un", "tan", ("col", 2))])
    want = O.projection(t1, [("un", "cos", ("col", 2)), ("un", "tan", ("col", 2))])
    assert want.cols[0].to_pylist() == want.cols[1].to_pylist() == [math.cos(x) for x in t1.cols[2].values]
    for c in (0, 1):
        for got, w in zip(out.cols[c].to_pylist(), want.cols[c].to_pylist()):
            assert abs(got - w) <= 4 * math.ulp(w)
    assert out.cols[0].to_pylist() == out.cols[1].to_pylist()  # bit-identical to each other: it IS cos
    # a few hundred random arguments, including large ones and NULLs
    rng = np.random.default_rng(3)
    b = O.Batch(["x"], [O.Col("f64", np.concatenate([rng.normal(0, 10, 300), rng.normal(0, 1e6, 100)]),
                               (rng.random(400) > 0.1).astype(np.uint8))])
    g = G.gpu_projection(b, [("un", "tan", ("col", 0))])
    w = O.projection(b, [("un", "tan", ("col", 0))])
    for got, want_ in zip(g.cols[0].to_pylist(), w.cols[0].to_pylist()):
        assert (got is None) == (want_ is None)
        if got is not None:
            assert abs(got - want_) <= 4 * math.ulp(want_)


def test_cast_panics_like_the_reference():  # cast.rs:45-87: every arm is todo!()
    import pyarrow as pa
    nq = G.nq
    b = O.Batch(["a"], [O.col("i64", [1, 2, 3])])
    cast = nq.PhysicalCastExpr.create(nq.ColumnExpr.try_create(None, 0), pa.float64())
    with pytest.raises(nq.NqeError) as e:
        cast.evaluate(G.to_arrow(b))
    assert e.value.kind == "Panic" and e.value.message == "not yet implemented"
    plan = nq.ProjectionPlan.create(G.scan(b), pa.schema([("a", pa.float64())]), [cast])
    with pytest.raises(nq.NqeError) as e:
        plan.execute()
    assert e.value.kind == "Panic"
    sel = nq.SelectionPlan.create(G.scan(b), nq.PhysicalBinaryExpr.create(cast, "Gt", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Float64(1.0))))
    with pytest.raises(nq.NqeError) as e:
        sel.execute()
    assert e.value.kind == "Panic"


def _strings(rng, n, null_frac):
    words = ["", "a", "ab", "abc", "b", "alice", "bob", "Zed", "\u00e9t\u00e9", "abcd" * 5]
    return O.col("utf8", [None if rng.random() < null_frac else words[int(rng.integers(0, len(words)))] + ("x" * int(rng.integers(0, 3)))
                          for _ in range(n)])


@pytest.mark.parametrize("op", ["Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq"])
def test_utf8_comparisons_in_expressions(op):
    """binary.rs:127-132: the *_dyn comparison kernels take Utf8 operands (bytewise order, NULL if either side is NULL):
    column vs column, column vs literal, literal vs column, a NULL literal, inside And, as a selection predicate (Utf8
    and numeric columns ride along) and as a projected Boolean."""
    rng = np.random.default_rng(len(op))
    n = 3000
    b = O.Batch(["s", "t", "x"], [_strings(rng, n, 0.1), _strings(rng, n, 0.0), rand_col(rng, "i64", n, 0.1)])
    cases = [("bin", op, ("col", 0), ("col", 1)), ("bin", op, ("col", 0), ("lit", "utf8", "ab")),
             ("bin", op, ("lit", "utf8", "b"), ("col", 1)), ("bin", op, ("col", 1), ("lit", "utf8", None)),
             ("bin", "And", ("bin", op, ("col", 0), ("col", 1)), ("bin", "Gt", ("col", 2), lit(0)))]
    for e in cases:
        same(G.gpu_projection(b, [e, ("col", 2)]), O.projection(b, [e, ("col", 2)]))
        same(G.gpu_selection(b, e), O.selection(b, e))


def test_utf8_selection_on_the_golden_table():  # `select * from t1 where name = 'alice'` through the physical operators
    t1 = golden_table("t1")
    got = G.gpu_selection(t1, ("bin", "Eq", ("col", 1), ("lit", "utf8", "alice")))
    want = O.selection(t1, ("bin", "Eq", ("col", 1), ("lit", "utf8", "alice")))
    same(got, want)
    assert got.num_rows == 1 and got.cols[1].to_pylist() == ["alice"]


def test_utf8_expression_errors():
    nq = G.nq
    b = O.Batch(["s", "t", "x"], [O.col("utf8", ["a", "b"]), O.col("utf8", ["a", "c"]), O.col("i64", [1, 2])])
    for e, kind in [(("bin", "Eq", ("col", 0), ("col", 2)), "IntervalError"), (("bin", "Plus", ("col", 0), ("col", 1)), "Panic"),
                    (("bin", "And", ("col", 0), ("col", 1)), "IntervalError"), (("bin", "Lt", ("col", 2), ("lit", "utf8", "a")), "IntervalError")]:
        with pytest.raises(nq.NqeError) as err:
            G.gpu_projection(b, [e])
        assert err.value.kind == kind, e
        with pytest.raises(O.OracleError) as oerr:
            O.projection(b, [e])
        assert oerr.value.kind == kind, e
    with pytest.raises(nq.NqeError) as err:
        G.gpu_projection(b, [("bin", "Eq", ("col", 0), ("col", 2))])
    assert err.value.message == "Cannot evaluate binary expression Eq with types Utf8 and Int64"


def _three_batches(rng, n, with_strings):
    """three ragged batches (17, n, 1 rows; offsets not multiples of 8 or 32) with NULLs, a Boolean and a Utf8 column,
    as oracle batches (raw values under NULL slots are part of the data: the join reads them, hash_join.rs:67)"""
    out = []
    for m in [17, n, 1]:
        cols = [O.Col("i64", rng.integers(0, 40, m).astype(np.int64), (rng.random(m) >= 0.1).astype(np.uint8)),
                O.Col("f64", np.round(rng.normal(0, 10, m), 3), (rng.random(m) >= 0.2).astype(np.uint8)),
                O.Col("bool", (rng.random(m) < 0.5).astype(np.uint8), (rng.random(m) >= 0.15).astype(np.uint8))]
        names = ["k", "v", "f"]
        if with_strings:
            cols.append(O.col("utf8", [None if rng.random() < 0.1 else "s" * int(rng.integers(0, 5)) + str(int(rng.integers(0, 9))) for _ in range(m)]))
            names.append("s")
        out.append(O.Batch(names, cols))
    return out


def _concat_oracle(batches):
    cols = []
    for i, c0 in enumerate(batches[0].cols):
        if c0.dtype == "utf8":
            vals = [v for b in batches for v in b.cols[i].to_pylist()]
            cols.append(O.col("utf8", vals))
        else:
            cols.append(O.Col(c0.dtype, np.concatenate([b.cols[i].values for b in batches]),
                              np.concatenate([b.cols[i].valid if b.cols[i].valid is not None else np.ones(b.num_rows, np.uint8) for b in batches])))
    return O.Batch(batches[0].names, cols)


@pytest.mark.parametrize("n", [0, 100, 5003])
def test_device_concat_equals_arrow_concat(n):  # concat_batches, hash_join.rs:258-273
    import pyarrow as pa
    nq = G.nq
    rng = np.random.default_rng(n)
    batches = [G.to_arrow(b) for b in _three_batches(rng, n, True)]
    parts = [nq.DeviceTable.from_arrow(b) for b in batches]
    got = nq.DeviceTable.concat(parts).to_arrow()
    want = pa.Table.from_batches(batches).combine_chunks().to_batches()[0]
    assert got.num_rows == want.num_rows
    for i in range(want.num_columns):
        assert got.column(i).to_pylist() == want.column(i).to_pylist(), want.schema.names[i]
        assert got.column(i).null_count == want.column(i).null_count


def test_multi_batch_inputs_through_join_and_group_by():
    """A 3-batch MemTable on the join's build side (hash_join.rs:131-132: concat_batches), on its probe side
    (:168-254: one output batch per probe batch, i.e. the concatenation) and under a GROUP BY (aggregate/mod.rs:143-144)."""
    import pyarrow as pa
    nq = G.nq
    rng = np.random.default_rng(8)
    ol, orr = _three_batches(rng, 300, False), _three_batches(rng, 2000, False)
    lb, rb = [G.to_arrow(b) for b in ol], [G.to_arrow(b) for b in orr]
    lsrc, rsrc = nq.MemTable.try_create(lb[0].schema, lb), nq.MemTable.try_create(rb[0].schema, rb)
    L, R = _concat_oracle(ol), _concat_oracle(orr)
    join = nq.HashJoin.create(nq.ScanPlan.create(lsrc, None), nq.ScanPlan.create(rsrc, None), [("k", "k")], "Inner")
    same(G.from_arrow(join.execute()[0]), O.hash_join_c(L, R, 0, 0))
    aggs = [("count", 1), ("sum", 1), ("avg", 1), ("min", 1), ("max", 1), ("min", 0)]
    plan = nq.PhysicalAggregatePlan.create([nq.ColumnExpr.try_create(None, 0)],
                                           [G._AGG[op].create(nq.ColumnExpr.try_create(None, ci)) for op, ci in aggs],
                                           nq.ScanPlan.create(rsrc, None))
    same(G.from_arrow(plan.execute()[0]), O.aggregate(R, ("col", 0), aggs), rel=SUM_REL, ordered=False, sort_cols=[5])


def test_golden_readme_groupby():  # README.md:105-111
    t1 = numeric_only(golden_table("t1"))
    out = G.gpu_aggregate(t1, ("bin", "Modulos", ("col", 0), lit(3)),
                          [("count", 0), ("sum", 1), ("sum", 2), ("avg", 2), ("max", 2), ("min", 2)])
    assert out.names == FX["readme_groupby"]["names"]
    assert_rows([list(r) for r in out.rows()], FX["readme_groupby"]["rows"], rel=SUM_REL, ordered=False)


def test_golden_readme_join_numeric_part():  # README.md:77-85 (ids and key columns; strings are a later row)
    emp, rank, dept = (numeric_only(golden_table(n)) for n in ("employee", "rank", "department"))
    j1 = G.gpu_join(emp, rank, "rank", "id")
    want1 = O.hash_join(emp, rank, "rank", "id")
    same(j1, want1)
    j2 = G.gpu_join(j1, dept, "department_id", "id")
    assert j2.cols[0].to_pylist() == [r[0] for r in FX["readme_join"]["rows"]]


# ---------------------------------------------------------------- filter / project
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 2047, 2048, 2049, 100_003])
@pytest.mark.parametrize("null_frac", [0.0, 0.2])
def test_filter_project_random(n, null_frac):
    rng = np.random.default_rng(n * 7 + int(null_frac * 10))
    b = O.Batch(["a", "b", "x", "f"], [rand_col(rng, "i64", n, null_frac), rand_col(rng, "i64", n, null_frac),
                                      rand_col(rng, "f64", n, null_frac), rand_col(rng, "bool", n, null_frac)])
    pred = ("bin", "Or", ("bin", "Lt", ("col", 0), lit(10)), ("bin", "And", ("col", 3), ("bin", "GtEq", ("col", 2), ("lit", "f64", 1.5))))
    exprs = [("col", 0), ("bin", "Plus", ("col", 1), lit(100)), ("bin", "Multiply", ("col", 2), ("lit", "f64", 0.5)),
             ("col", 3), ("bin", "Minus", ("bin", "Multiply", ("col", 0), ("col", 1)), ("bin", "Plus", ("col", 1), lit(7))),
             ("bin", "NotEq", ("col", 0), ("col", 1))]
    want = O.projection(O.selection(b, pred), exprs)
    got = G.gpu_projection(b, exprs, pred=pred)
    same(got, want)
    # bare SelectionPlan (all columns pass through) and bare ProjectionPlan
    same(G.gpu_selection(b, pred), O.selection(b, pred))
    same(G.gpu_projection(b, exprs), O.projection(b, exprs))


def _fp_arrow(n, seed, k):
    """Numeric NULL-free table + `a < k` / (a, b + 100, x * 0.5, f) through the fused call; numpy-side expectation."""
    import pyarrow as pa
    from importlib import import_module
    pp = import_module("naive-query-engine_b200.physical_plan")
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1000, n).astype(np.int64)
    b = rng.integers(-100, 100, n).astype(np.int64)
    x = rng.normal(0, 10, n)
    f = rng.integers(0, 2, n).astype(bool)
    rb = pa.RecordBatch.from_arrays([pa.array(a), pa.array(b), pa.array(x), pa.array(f)], names=["a", "b", "x", "f"])
    src = G.nq.ScanPlan.create(G.nq.MemTable.try_create(rb.schema, [rb]), None).execute_device()
    pred = G.expr(("bin", "Lt", ("col", 0), lit(k)))
    exprs = [G.expr(e) for e in [("col", 0), ("bin", "Plus", ("col", 1), lit(100)),
                                 ("bin", "Multiply", ("col", 2), ("lit", "f64", 0.5)), ("col", 3)]]
    got = pp._filter_project(src, pred, exprs, ["a", "b + 100", "x * 0.5", "f"]).to_arrow()
    m = a < k
    return got, [a[m], b[m] + 100, x[m] * 0.5, f[m]]


@pytest.mark.parametrize("n", [1024 * 37, 1024 * 37 + 1, 3_000_017])
@pytest.mark.parametrize("k", [0, 7, 500, 993, 1000])
def test_filter_project_many_tiles(n, k):
    """Multi-tile look-back chain, every selectivity, ragged and exact tile boundaries, Boolean column staged as a bitmap."""
    got, want = _fp_arrow(n, n + k, k)
    assert got.num_rows == len(want[0])
    for i in range(4):
        assert np.array_equal(got.column(i).to_numpy(zero_copy_only=False), want[i]), f"column {i}"


@pytest.mark.parametrize("n", [2048 * 300, 1_000_003])
@pytest.mark.parametrize("null_frac", [0.02, 0.5])
def test_filter_project_many_tiles_nullable(n, null_frac):
    """NULL-aware two-ring kernel over hundreds of tiles: validity bitmaps staged next to the values, Kleene OR,
    rows whose predicate is NULL kept as all-NULL rows (selection.rs:46), divide with NULL/dropped zero divisors."""
    from importlib import import_module
    pp = import_module("naive-query-engine_b200.physical_plan")
    rng = np.random.default_rng(n + int(null_frac * 100))
    b = O.Batch(["a", "b", "x", "f"], [rand_col(rng, "i64", n, null_frac), rand_col(rng, "i64", n, null_frac, lo=1, hi=60),
                                      rand_col(rng, "f64", n, null_frac), rand_col(rng, "bool", n, null_frac)])
    pred = ("bin", "Or", ("bin", "Lt", ("col", 0), lit(10)), ("bin", "And", ("col", 3), ("bin", "GtEq", ("col", 2), ("lit", "f64", 1.5))))
    exprs = [("col", 0), ("bin", "Plus", ("col", 1), lit(100)), ("bin", "Multiply", ("col", 2), ("lit", "f64", 0.5)), ("col", 3),
             ("bin", "Divide", ("col", 0), ("col", 1)), ("bin", "LtEq", ("col", 0), ("col", 1)), ("lit", "i64", None)]
    want = O.projection(O.selection(b, pred), exprs)
    src = G.scan(b).execute_device()
    got = pp._filter_project(src, G.expr(pred), [G.expr(e) for e in exprs], [f"o{i}" for i in range(len(exprs))]).to_arrow()
    assert got.num_rows == want.num_rows
    for i, w in enumerate(want.cols):
        a = got.column(i)
        gv = np.ones(len(a), dtype=np.uint8) if a.null_count == 0 else np.asarray(a.is_valid()).astype(np.uint8)
        wv = np.ones(len(a), dtype=np.uint8) if w.valid is None else w.valid
        assert np.array_equal(gv, wv), f"validity of column {i}"
        g = a.fill_null(False if w.dtype == "bool" else 0).to_numpy(zero_copy_only=False)
        m = wv != 0
        assert np.array_equal(np.asarray(g)[m].astype(w.values.dtype), w.values[m]), f"values of column {i}"


@pytest.mark.parametrize("ncols", [6, 12, 16])
def test_filter_project_wide_table(ncols):
    """Many referenced columns: the two-ring kernel shrinks its tile (K = 8 -> 4 -> 2 -> 1) so that both rings still
    leave room for two CTAs per SM; every column is projected, half of them through an expression."""
    import pyarrow as pa
    from importlib import import_module
    pp = import_module("naive-query-engine_b200.physical_plan")
    rng = np.random.default_rng(ncols)
    n = 150_001
    cols = [rng.integers(-1000, 1000, n).astype(np.int64) for _ in range(ncols)]
    rb = pa.RecordBatch.from_arrays([pa.array(c) for c in cols], names=[f"c{i}" for i in range(ncols)])
    src = G.nq.ScanPlan.create(G.nq.MemTable.try_create(rb.schema, [rb]), None).execute_device()
    pred = G.expr(("bin", "Lt", ("col", 0), lit(300)))
    exprs = [G.expr(("col", i) if i % 2 == 0 else ("bin", "Plus", ("col", i), lit(i))) for i in range(ncols)]
    got = pp._filter_project(src, pred, exprs, [f"o{i}" for i in range(ncols)]).to_arrow()
    m = cols[0] < 300
    assert got.num_rows == int(m.sum())
    for i in range(ncols):
        want = cols[i][m] if i % 2 == 0 else cols[i][m] + i
        assert np.array_equal(got.column(i).to_numpy(), want), f"column {i}"


def test_filter_project_full_size_config():
    """BASELINE configs[1] at full size (1e8 rows generated in HBM): exact equality with the host-side generator."""
    import torch
    from importlib import import_module
    synth = import_module("naive-query-engine_b200.synth")
    pp = import_module("naive-query-engine_b200.physical_plan")
    ctx = G.nq.Context.default()
    n = 100_000_000
    bufs = []
    for spec in synth.FILTER_TABLE:
        t = torch.empty(n, dtype=torch.int64, device="cuda")
        synth.device_column(ctx, spec, 0, n, t.data_ptr())
        bufs.append(t)
    ctx.sync()
    tbl = G.nq.DeviceTable.from_device_pointers(ctx, ["id", "age", "score"], [2, 2, 4], [b.data_ptr() for b in bufs], n, keepalive=bufs)
    pred = G.expr(("bin", "Lt", ("col", 0), lit(500)))
    out = pp._filter_project(tbl, pred, [G.expr(("col", 0)), G.expr(("bin", "Plus", ("col", 1), lit(100)))], ["id", "age + 100"])
    rows = out.num_rows
    desc = [out.column_desc(i) for i in range(2)]
    got = [torch.as_tensor(_CAI(d.values, rows), device="cuda").cpu().numpy() for d in desc]
    ids = synth.mod_i64(42, 0, n, 1000)
    m = ids < 500
    assert rows == int(m.sum())
    assert np.array_equal(got[0], ids[m])
    del ids
    age = synth.mod_i64(43, 0, n, 100)
    assert np.array_equal(got[1], age[m] + 100)
    out.free()
    tbl.free()


@pytest.mark.parametrize("n", [20_000_017, 100_003])
def test_filter_project_host_pipeline(n):
    """nqe_filter_project_host: pinned host columns in, pinned host buffers out; the large case runs the chunked
    three-stream pipeline (only `id` and `age` are uploaded), the small one the plain upload/operator/download path."""
    import torch
    from importlib import import_module
    pp = import_module("naive-query-engine_b200.physical_plan")
    synth = import_module("naive-query-engine_b200.synth")
    ctx = G.nq.Context.default()
    ids, age, score = synth.mod_i64(42, 0, n, 1000), synth.mod_i64(43, 0, n, 100), synth.unif_f64(44, 0, n, 100.0)
    host = [torch.from_numpy(a.view(np.int64)).pin_memory() for a in (ids, age, score)]
    outs = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(3)]
    pred = G.expr(("bin", "Lt", ("col", 0), lit(500)))
    exprs = [G.expr(("col", 0)), G.expr(("bin", "Plus", ("col", 1), lit(100))), G.expr(("bin", "Multiply", ("col", 2), ("lit", "f64", 2.0)))]
    rows, dts = pp.filter_project_host(ctx, ["id", "age", "score"], [2, 2, 4], [h.data_ptr() for h in host], n, pred, exprs,
                                       [o.data_ptr() for o in outs], n)
    m = ids < 500
    assert rows == int(m.sum()) and dts == [2, 2, 4]
    assert np.array_equal(outs[0].numpy()[:rows], ids[m])
    assert np.array_equal(outs[1].numpy()[:rows], age[m] + 100)
    assert np.array_equal(outs[2].numpy()[:rows].view(np.float64), score[m] * 2.0)


class _CAI:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


@pytest.mark.parametrize("sel", ["none", "all"])
def test_filter_all_or_nothing(sel):
    rng = np.random.default_rng(5)
    n = 10_000
    b = O.Batch(["a", "x"], [rand_col(rng, "i64", n), rand_col(rng, "f64", n)])
    pred = ("bin", "Lt", ("col", 0), lit(-1000 if sel == "none" else 1000))
    same(G.gpu_selection(b, pred), O.selection(b, pred))


OPS_CMP = ["Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq"]
OPS_ARITH = ["Plus", "Minus", "Multiply", "Divide", "Modulos"]


@pytest.mark.parametrize("dtype", ["i64", "u64", "f64"])
def test_every_operator(dtype):
    rng = np.random.default_rng(11)
    n = 5000
    a = rand_col(rng, dtype, n, 0.1)
    bvals = rand_col(rng, dtype, n, 0.1)
    # non-zero divisors (zero divisors are covered by the error test)
    if dtype == "f64":
        bvals.values[bvals.values == 0.0] = 1.25
        a.values[:5] = [float("nan"), float("inf"), -float("inf"), 0.0, -0.0]
    else:
        bvals.values[bvals.values == 0] = 3
    b = O.Batch(["a", "b"], [a, bvals])
    exprs = [("bin", op, ("col", 0), ("col", 1)) for op in OPS_CMP + OPS_ARITH]
    same(G.gpu_projection(b, exprs), O.projection(b, exprs))
    edge = {"i64": [2**63 - 1, -2**63, -1, 0, 1], "u64": [2**64 - 1, 0, 1, 2**63, 5], "f64": [1e308, -1e308, 5e-324, 0.1, 3.0]}[dtype]
    e = O.Batch(["a", "b"], [O.col(dtype, edge), O.col(dtype, list(reversed(edge)))])
    exprs = [("bin", op, ("col", 0), ("col", 1)) for op in OPS_CMP + ["Plus", "Minus", "Multiply"]]
    same(G.gpu_projection(e, exprs), O.projection(e, exprs))


def test_kleene_and_bool_compare():
    T, F, N = True, False, None
    b = O.Batch(["a", "b"], [O.col("bool", [T, T, T, F, F, F, N, N, N]), O.col("bool", [T, F, N, T, F, N, T, F, N])])
    exprs = [("bin", "And", ("col", 0), ("col", 1)), ("bin", "Or", ("col", 0), ("col", 1)),
             ("bin", "Eq", ("col", 0), ("col", 1)), ("bin", "Lt", ("col", 0), ("col", 1)),
             ("bin", "And", ("col", 0), ("lit", "bool", None)), ("bin", "Or", ("col", 0), ("lit", "bool", True))]
    same(G.gpu_projection(b, exprs), O.projection(b, exprs))


def test_null_predicate_rows_kept_as_null_rows():  # selection.rs:46
    b = O.Batch(["a", "b"], [O.col("i64", [1, None, 3, 4]), O.col("f64", [1.5, 2.5, None, 4.5])])
    pred = ("bin", "Gt", ("col", 0), lit(2))
    got = G.gpu_selection(b, pred)
    assert got.rows() == [(None, None), (3, None), (4, 4.5)]
    exprs = [("bin", "Plus", ("col", 0), lit(1)), ("lit", "i64", 5)]
    same(G.gpu_projection(b, exprs, pred=pred), O.projection(O.selection(b, pred), exprs))


def test_deep_expression_uses_stack():
    rng = np.random.default_rng(3)
    n = 3000
    b = O.Batch(["a", "b", "c"], [rand_col(rng, "i64", n), rand_col(rng, "i64", n, 0.1), rand_col(rng, "i64", n)])
    e = ("bin", "Minus", lit(5), ("bin", "Multiply", ("bin", "Plus", ("col", 0), ("col", 1)),
                                   ("bin", "Minus", ("col", 2), ("bin", "Plus", ("col", 0), lit(3)))))
    same(G.gpu_projection(b, [e]), O.projection(b, [e]))


@pytest.mark.parametrize("case", ["interval", "div0_int", "div0_float", "mod0", "overflow", "and_on_int", "arith_on_bool",
                                  "unary_on_int", "filtered_div0_ok"])
def test_error_behaviour(case):
    import nqe_b200 as nq
    b = O.Batch(["a", "z", "f", "t"], [O.col("i64", [1, 2, -2**63]), O.col("i64", [1, 0, -1]),
                                      O.col("f64", [1.0, 0.0, 2.0]), O.col("bool", [True, False, True])])
    cases = {
        "interval": (("bin", "Lt", ("col", 0), ("lit", "f64", 9.5)), "IntervalError"),
        "div0_int": (("bin", "Divide", ("col", 0), ("col", 1)), "ArrowError(DivideByZero)"),
        "div0_float": (("bin", "Divide", ("col", 2), ("col", 2)), "ArrowError(DivideByZero)"),
        "mod0": (("bin", "Modulos", ("col", 0), ("col", 1)), "ArrowError(DivideByZero)"),
        "and_on_int": (("bin", "And", ("col", 0), ("col", 1)), "IntervalError"),
        "arith_on_bool": (("bin", "Plus", ("col", 3), ("col", 3)), "Panic"),
        "unary_on_int": (("un", "abs", ("col", 0)), "Panic"),
    }
    if case == "overflow":
        b2 = O.Batch(["a", "z"], [O.col("i64", [-2**63]), O.col("i64", [-1])])
        with pytest.raises(nq.NqeError) as e:
            G.gpu_projection(b2, [("bin", "Divide", ("col", 0), ("col", 1))])
        assert e.value.kind == "Panic"
        with pytest.raises(O.OracleError) as oe:
            O.projection(b2, [("bin", "Divide", ("col", 0), ("col", 1))])
        assert oe.value.kind == "Panic"
        return
    if case == "filtered_div0_ok":
        # the zero divisor sits in a row the selection drops: the projection never sees it
        pred = ("bin", "NotEq", ("col", 1), lit(0))
        ex = [("bin", "Divide", ("col", 0), ("col", 1))]
        b3 = O.Batch(["a", "z"], [O.col("i64", [10, 20, 30]), O.col("i64", [2, 0, 5])])
        same(G.gpu_projection(b3, ex, pred=pred), O.projection(O.selection(b3, pred), ex))
        return
    ex, kind = cases[case]
    with pytest.raises(nq.NqeError) as e:
        G.gpu_projection(b, [ex])
    assert e.value.kind == kind
    with pytest.raises(O.OracleError) as oe:
        O.projection(b, [ex])
    assert oe.value.kind == kind
    if case == "interval":
        assert e.value.message == oe.value.msg == "Cannot evaluate binary expression Lt with types Int64 and Float64"


# ---------------------------------------------------------------- hash join
@pytest.mark.parametrize("nl,nr,keyspace", [(0, 10, 5), (10, 0, 5), (1, 1, 1), (1000, 5000, 1000), (5000, 1000, 100),
                                            (3000, 3000, 10**9), (20000, 100_000, 20000)])
def test_hash_join_random(nl, nr, keyspace):
    rng = np.random.default_rng(nl + nr)
    l = O.Batch(["k", "a", "lb"], [O.Col("i64", rng.integers(0, keyspace, nl)), rand_col(rng, "i64", nl, 0.1),
                                  rand_col(rng, "bool", nl)])
    r = O.Batch(["fk", "b"], [O.Col("i64", rng.integers(0, keyspace, nr)), rand_col(rng, "f64", nr, 0.1)])
    want = O.hash_join_c(l, r, 0, 0)
    got = G.gpu_join(l, r, "k", "fk")
    same(got, want)  # identical order: probe-row-major, build rows ascending


@pytest.mark.parametrize("n_pay,poison", [(1, False), (1, True), (2, False), (3, False)])
def test_hash_join_partitioned_probe_large(n_pay, poison):
    """Build table larger than the L2 budget -> the partitioned (radix) probe path; result must equal the
    probe-row-major order of the reference (hash_join.rs:80-103).  n_pay == 1: the payload rides in the slots' row word
    (poison: some payloads equal the free-slot marker, so the build falls back to row numbers); 2, 3: thin slots + gather."""
    import pyarrow as pa
    rng = np.random.default_rng(77 + n_pay)
    nl, nr = 2_000_000, 5_000_017
    keys = rng.permutation(np.arange(1, 3 * nl + 1, dtype=np.int64))[:nl] * 1_000_003  # unique, sparse
    pays = [rng.integers(-1 << 40, 1 << 40, nl).astype(np.int64) for _ in range(n_pay)]
    if poison:  # a single payload column rides in the slots' row word -- unless a value equals the free-slot marker
        pays[0][rng.integers(0, nl, 50)] = -1
    fk = np.where(rng.random(nr) < 0.7, keys[rng.integers(0, nl, nr)], rng.integers(1, 1 << 50, nr) * 2 + 7).astype(np.int64)
    b = rng.normal(0, 10, nr)
    L = pa.RecordBatch.from_arrays([pa.array(keys)] + [pa.array(p_) for p_ in pays], names=["k"] + [f"p{i}" for i in range(n_pay)])
    R = pa.RecordBatch.from_arrays([pa.array(fk), pa.array(b)], names=["fk", "b"])
    nq = G.nq
    join = nq.HashJoin.create(nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None),
                              nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None), [("k", "fk")], "Inner")
    got = join.execute()[0]
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    idx = np.searchsorted(sk, fk)
    idx[idx >= nl] = nl - 1
    m = sk[idx] == fk
    brow = order[idx[m]]
    assert got.num_rows == int(m.sum())
    assert np.array_equal(got.column(0).to_numpy(), fk[m])
    for i in range(n_pay):
        assert np.array_equal(got.column(1 + i).to_numpy(), pays[i][brow]), f"payload {i}"
    assert np.array_equal(got.column(1 + n_pay).to_numpy(), fk[m])
    assert np.array_equal(got.column(2 + n_pay).to_numpy(), b[m])


@pytest.mark.parametrize("group_side", ["build", "probe"])
def test_join_aggregate_partitioned_large(group_side):
    """Fused join -> group-by with a build table larger than the L2 budget; 30 % of the probe rows find no match.
    (With NQE_JOINAGG_PART=1 in the environment this exercises the radix-partitioned variant of the fused path.)"""
    import ctypes as C
    import pyarrow as pa
    rng = np.random.default_rng(5)
    nl, nr, g = 2_000_000, 5_000_011, 777
    keys = rng.permutation(np.arange(1, 3 * nl + 1, dtype=np.int64))[:nl] * 1_000_003
    a = rng.integers(0, g, nl).astype(np.int64)
    pick = rng.integers(0, nl, nr)
    hit = rng.random(nr) < 0.7
    fk = np.where(hit, keys[pick], rng.integers(1, 1 << 50, nr) * 2 + 7).astype(np.int64)
    b = np.round(rng.normal(0, 10, nr), 3)
    c = rng.integers(0, g, nr).astype(np.int64)
    L = pa.RecordBatch.from_arrays([pa.array(keys), pa.array(a)], names=["k", "a"])
    R = pa.RecordBatch.from_arrays([pa.array(fk), pa.array(b), pa.array(c)], names=["fk", "b", "c"])
    nq = G.nq
    lt = nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None).execute_device()
    rt = nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None).execute_device()
    gcol = 1 if group_side == "build" else 4  # joined schema: k, a, fk, b, c
    aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, col) for o, col in [(3, gcol), (0, 3), (1, 3), (3, 3), (4, 3)]])
    h = C.c_void_p()
    ctx = lt.ctx
    ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, gcol, aggs, 5, C.byref(h)))
    out = nq.DeviceTable(ctx, h, ["key", "count", "sum", "min", "max"]).to_arrow()
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    idx = np.minimum(np.searchsorted(sk, fk), nl - 1)
    m = sk[idx] == fk
    grp = a[order[idx[m]]] if group_side == "build" else c[m]
    bv = b[m]
    key = out.column(0).to_numpy().astype(np.int64)
    o = np.argsort(key)
    present = np.unique(grp)
    assert np.array_equal(key[o], present)
    cnt = np.bincount(grp, minlength=g)[present]
    assert np.array_equal(out.column(1).to_numpy()[o], cnt.astype(np.uint64))
    assert np.allclose(out.column(2).to_numpy()[o], np.bincount(grp, weights=bv, minlength=g)[present], rtol=SUM_REL, atol=1e-6)
    mn = np.full(g, np.inf); mx = np.full(g, -np.inf)
    np.minimum.at(mn, grp, bv); np.maximum.at(mx, grp, bv)
    assert np.array_equal(out.column(3).to_numpy()[o], mn[present]) and np.array_equal(out.column(4).to_numpy()[o], mx[present])


def _group_rows(cnt, sm, mn, mx):
    """sortable (count, min, max, sum) rows -- the aggregate output carries no key column, so groups are matched through
    their (count, min, max) triples"""
    o = np.lexsort((mx, mn, cnt))
    return cnt[o], mn[o], mx[o], sm[o]


@pytest.mark.parametrize("nl,g", [(2_000_000, 60_000), (5_000_000, 150_000)])
def test_join_aggregate_paged_large(nl, g):
    """The paged fused join -> group-by (hash_join.cu: split by slot range, probe + re-split by group hash, shared-memory
    aggregation): unique build keys, group key from the build side, every aggregate over one probe-side column; 30 % of
    the probe rows find no match; the group keys include i64::MIN (the group tables' free-slot marker)."""
    import ctypes as C
    import pyarrow as pa
    rng = np.random.default_rng(nl // 1000 + g)
    nr = 6_000_011
    keys = rng.permutation(np.arange(1, 3 * nl + 1, dtype=np.int64))[:nl] * 1_000_003
    a = rng.integers(0, g, nl).astype(np.int64) * 7919 - 5
    a[a == -5] = np.iinfo(np.int64).min
    fk = np.where(rng.random(nr) < 0.7, keys[rng.integers(0, nl, nr)], rng.integers(1, 1 << 50, nr) * 2 + 7).astype(np.int64)
    b = np.round(rng.normal(0, 1000, nr), 6)
    b[rng.integers(0, nr, 20)] = np.nan
    b[rng.integers(0, nr, 20)] = np.inf
    L = pa.RecordBatch.from_arrays([pa.array(keys), pa.array(a)], names=["k", "a"])
    R = pa.RecordBatch.from_arrays([pa.array(fk), pa.array(b)], names=["fk", "b"])
    nq = G.nq
    lt = nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None).execute_device()
    rt = nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None).execute_device()
    aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, 3) for o in (0, 1, 2, 3, 4)])  # joined schema: k, a, fk, b
    h = C.c_void_p()
    ctx = lt.ctx
    ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, 1, aggs, 5, C.byref(h)))
    out = nq.DeviceTable(ctx, h, ["count", "sum", "avg", "min", "max"]).to_arrow()
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    idx = np.minimum(np.searchsorted(sk, fk), nl - 1)
    m = sk[idx] == fk
    grp, bv = a[order[idx[m]]], b[m]
    uk, inv = np.unique(grp, return_inverse=True)
    ng = len(uk)
    assert out.num_rows == ng
    cnt = np.bincount(inv, minlength=ng)
    so = np.argsort(inv, kind="stable")
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    vs = bv[so]
    has_nan = np.bincount(inv, weights=np.isnan(bv), minlength=ng) > 0
    mx = np.where(has_nan, np.nan, np.maximum.reduceat(np.where(np.isnan(vs), -np.inf, vs), starts))
    mn = np.minimum.reduceat(np.where(np.isnan(vs), np.inf, vs), starts)
    with np.errstate(invalid="ignore"):
        sm = np.add.reduceat(vs, starts)
    key = lambda x: np.nan_to_num(x, nan=1e308, posinf=1e307, neginf=-1e307)
    got = [out.column(i).to_numpy(zero_copy_only=False) for i in range(5)]
    gc, gmn, gmx, gsm = _group_rows(got[0].astype(np.int64), got[1], key(got[3]), key(got[4]))
    wc, wmn, wmx, wsm = _group_rows(cnt, sm, key(mn), key(mx))
    assert np.array_equal(gc, wc) and np.array_equal(gmn, wmn) and np.array_equal(gmx, wmx)  # count, min, max: exact
    fin = np.isfinite(wsm)
    assert np.array_equal(np.isnan(gsm), np.isnan(wsm))
    scale = np.add.reduceat(np.abs(np.where(np.isfinite(vs), vs, 0.0)), starts)[np.lexsort((key(mx), key(mn), cnt))]
    assert np.all(np.abs(gsm[fin] - wsm[fin]) <= SUM_REL * np.maximum(scale[fin], 1.0))
    lt.free()
    rt.free()


def test_hash_join_partitioned_probe_nullable_probe_key():
    """Partition-size join whose PROBE key column has NULL slots: key validity is ignored (hash_join.rs:67,86), a NULL
    probe key whose raw value matches joins, the build key output is the build side's (valid) value and only the probe
    key output carries the NULL -- the same answer the direct probe path gives on small inputs."""
    import pyarrow as pa
    rng = np.random.default_rng(123)
    nl, nr = 2_000_000, 4_500_003
    keys = rng.permutation(np.arange(1, 3 * nl + 1, dtype=np.int64))[:nl] * 1_000_003
    pay = rng.integers(-1 << 40, 1 << 40, nl).astype(np.int64)
    fk = np.where(rng.random(nr) < 0.8, keys[rng.integers(0, nl, nr)], rng.integers(1, 1 << 50, nr) * 2 + 7).astype(np.int64)
    null = rng.random(nr) < 0.1
    b = rng.normal(0, 10, nr)
    L = pa.RecordBatch.from_arrays([pa.array(keys), pa.array(pay)], names=["k", "p"])
    R = pa.RecordBatch.from_arrays([pa.array(fk, mask=null), pa.array(b)], names=["fk", "b"])
    nq = G.nq
    join = nq.HashJoin.create(nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None),
                              nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None), [("k", "fk")], "Inner")
    got = join.execute()[0]
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    idx = np.minimum(np.searchsorted(sk, fk), nl - 1)
    m = sk[idx] == fk
    assert got.num_rows == int(m.sum())
    assert got.column(0).null_count == 0 and np.array_equal(got.column(0).to_numpy(), fk[m])
    assert np.array_equal(got.column(1).to_numpy(), pay[order[idx[m]]])
    assert got.column(2).null_count == int((null & m).sum())
    assert np.array_equal(np.asarray(got.column(2).is_null()), null[m])
    assert np.array_equal(got.column(3).to_numpy(), b[m])


def _np_unique_key_join(keys, fk):
    """(match mask over the probe rows, build row of every match) for unique build keys"""
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    idx = np.minimum(np.searchsorted(sk, fk), len(keys) - 1)
    m = sk[idx] == fk
    return m, order[idx[m]]


def _dense_keys(rng, nl, lo, holes=True):
    """unique keys out of [lo, lo + range): a permutation with (holes) about a third of the range unused"""
    span = nl + nl // 2 if holes else nl
    return (rng.permutation(span)[:nl] + lo).astype(np.int64)


@pytest.mark.parametrize("n_pay,variant", [(1, "plain"), (1, "poison"), (1, "negative"), (2, "plain"), (2, "nullable"), (0, "plain")])
def test_hash_join_direct_table(n_pay, variant):
    """Dense unique build keys -> the direct table (hash_join.cu JoinTable::direct): slot = key - min, one 8-byte read
    per probe row, probed in place by the general kernel.  Probe keys hit present keys, holes inside the range and keys
    below / above it.  n_pay == 1: the payload rides in the table (poison: a payload equals the free-slot marker, so the
    host falls back to the hashed table); 2: row numbers + gathers, with a nullable and a Boolean build column."""
    import pyarrow as pa
    rng = np.random.default_rng(1000 + n_pay + len(variant))
    nl, nr = 300_000, 2_000_003
    lo = -123_456_789 if variant == "negative" else 5_000
    keys = _dense_keys(rng, nl, lo)
    fk = rng.integers(lo - nl // 4, lo + 2 * nl, nr).astype(np.int64)  # present keys, holes, out of range on both sides
    b = rng.normal(0, 10, nr)
    pays = [rng.integers(-1 << 40, 1 << 40, nl).astype(np.int64) for _ in range(n_pay)]
    if variant == "poison":
        pays[0][rng.integers(0, nl, 20)] = -1
    arrays = [pa.array(keys)] + [pa.array(p_) for p_ in pays]
    names = ["k"] + [f"p{i}" for i in range(n_pay)]
    pmask = flag = None
    if variant == "nullable":
        pmask = rng.random(nl) < 0.2
        arrays[1] = pa.array(pays[0], mask=pmask)
        flag = rng.random(nl) < 0.5
        arrays.append(pa.array(flag))
        names.append("flag")
    L = pa.RecordBatch.from_arrays(arrays, names=names)
    R = pa.RecordBatch.from_arrays([pa.array(fk), pa.array(b)], names=["fk", "b"])
    nq = G.nq
    join = nq.HashJoin.create(nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None),
                              nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None), [("k", "fk")], "Inner")
    got = join.execute()[0]
    m, brow = _np_unique_key_join(keys, fk)
    assert 0 < int(m.sum()) < nr and got.num_rows == int(m.sum())
    assert np.array_equal(got.column(0).to_numpy(), fk[m])
    for i in range(n_pay):
        col = got.column(1 + i)
        if i == 0 and pmask is not None:
            assert np.array_equal(np.asarray(col.is_null()), pmask[brow])
            ok = ~pmask[brow]
            assert np.array_equal(col.to_numpy(zero_copy_only=False)[ok].astype(np.int64), pays[0][brow][ok])
        else:
            assert np.array_equal(col.to_numpy(), pays[i][brow]), f"payload {i}"
    nleft = len(names)
    if flag is not None:
        assert np.array_equal(got.column(nleft - 1).to_numpy(zero_copy_only=False), flag[brow])
    assert np.array_equal(got.column(nleft).to_numpy(), fk[m])
    assert np.array_equal(got.column(nleft + 1).to_numpy(), b[m])


@pytest.mark.parametrize("nr,g,dense_groups", [(6_000_011, 60_000, True), (6_000_011, 150_000, False), (300_007, 5_000, True)])
def test_join_aggregate_direct_table(nr, g, dense_groups):
    """Fused join -> group-by over a direct table (dense unique build keys): the probe rows are read in place, become
    (group key, value) and go straight into the group-by's partitions (ja_direct_scatter_kernel); the small probe side
    takes the single-pass kernel over the same table.  A quarter of the probe keys miss (holes and out of range)."""
    import ctypes as C
    import pyarrow as pa
    rng = np.random.default_rng(nr % 1000 + g)
    nl = 2_000_000
    keys = _dense_keys(rng, nl, -7_000)
    a = rng.integers(0, g, nl).astype(np.int64)
    if not dense_groups:
        a = a * 7919 - 5
        a[a == -5] = np.iinfo(np.int64).min
    fk = rng.integers(-7_000 - nl // 8, -7_000 + nl + nl // 2 + nl // 8, nr).astype(np.int64)
    b = np.round(rng.normal(0, 1000, nr), 6)
    L = pa.RecordBatch.from_arrays([pa.array(keys), pa.array(a)], names=["k", "a"])
    R = pa.RecordBatch.from_arrays([pa.array(fk), pa.array(b)], names=["fk", "b"])
    nq = G.nq
    lt = nq.ScanPlan.create(nq.MemTable.try_create(L.schema, [L]), None).execute_device()
    rt = nq.ScanPlan.create(nq.MemTable.try_create(R.schema, [R]), None).execute_device()
    aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, c) for o, c in [(5, 0), (0, 3), (1, 3), (3, 3), (4, 3)]])  # key, count, sum, min, max of b
    h = C.c_void_p()
    ctx = lt.ctx
    ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, lt.h, rt.h, 0, 0, 1, aggs, 5, C.byref(h)))
    out = nq.DeviceTable(ctx, h, ["key", "count", "sum", "min", "max"]).to_arrow()
    m, brow = _np_unique_key_join(keys, fk)
    grp, bv = a[brow], b[m]
    uk, inv = np.unique(grp, return_inverse=True)
    key = out.column(0).to_numpy().astype(np.int64)
    o = np.argsort(key)
    assert np.array_equal(key[o], uk)
    cnt = np.bincount(inv, minlength=len(uk))
    assert np.array_equal(out.column(1).to_numpy()[o], cnt.astype(np.uint64))
    scale = np.bincount(inv, weights=np.abs(bv), minlength=len(uk))
    assert np.all(np.abs(out.column(2).to_numpy()[o] - np.bincount(inv, weights=bv, minlength=len(uk))) <= SUM_REL * np.maximum(scale, 1.0))
    mn = np.full(len(uk), np.inf); mx = np.full(len(uk), -np.inf)
    np.minimum.at(mn, inv, bv); np.maximum.at(mx, inv, bv)
    assert np.array_equal(out.column(3).to_numpy()[o], mn) and np.array_equal(out.column(4).to_numpy()[o], mx)
    lt.free()
    rt.free()


def test_hash_join_unique_build_keys_and_u64():
    rng = np.random.default_rng(9)
    nl, nr = 10_000, 50_000
    l = O.Batch(["k", "a"], [O.Col("u64", rng.permutation(nl).astype(np.uint64)), rand_col(rng, "i64", nl)])
    r = O.Batch(["fk", "b"], [O.Col("u64", rng.integers(0, 2 * nl, nr).astype(np.uint64)), rand_col(rng, "f64", nr)])
    same(G.gpu_join(l, r, "k", "fk"), O.hash_join_c(l, r, 0, 0))


def test_hash_join_ignores_key_validity():  # hash_join.rs:67,86
    l = O.Batch(["k", "a"], [O.col("i64", [5, 7, 5, 0], valid=[1, 1, 1, 0]), O.col("i64", [10, 11, 12, 13])])
    r = O.Batch(["fk", "b"], [O.col("i64", [7, 5, 0, 9]), O.col("f64", [0.5, 1.5, 2.5, 3.5])])
    got = G.gpu_join(l, r, "k", "fk")
    assert got.rows() == [(7, 11, 7, 0.5), (5, 10, 5, 1.5), (5, 12, 5, 1.5), (None, 13, 0, 2.5)]


def test_hash_join_errors():
    import nqe_b200 as nq
    l = O.Batch(["k"], [O.col("f64", [1.0])])
    r = O.Batch(["fk"], [O.col("f64", [1.0])])
    with pytest.raises(nq.NqeError) as e:
        G.gpu_join(l, r, "k", "fk")
    assert e.value.kind == "NotImplemented"  # hash_join.rs:161
    li = O.Batch(["k"], [O.col("i64", [1])])
    with pytest.raises(nq.NqeError) as e:
        nq.HashJoin.create(G.scan(li), G.scan(li), [], "Inner").execute()
    assert e.value.kind == "PlanError" and e.value.message == "Inner Join on Conditions can't not be empty"


# ---------------------------------------------------------------- aggregates
ALL_AGGS = [("count", 1), ("sum", 1), ("avg", 1), ("min", 1), ("max", 1)]


@pytest.mark.parametrize("n,groups", [(0, 5), (1, 1), (1000, 3), (50_000, 1000), (200_000, 150_000)])
@pytest.mark.parametrize("vtype", ["f64", "i64"])
def test_group_by_random(n, groups, vtype):
    rng = np.random.default_rng(n + groups)
    k = O.Col("i64", rng.integers(-groups // 2, groups // 2 + 1, n), (rng.random(n) > 0.05).astype(np.uint8))
    v = rand_col(rng, vtype, n, 0.1, lo=-10**6, hi=10**6)
    b = O.Batch(["k", "v"], [k, v])
    want = O.aggregate(b, ("col", 0), ALL_AGGS + [("min", 0)])
    got = G.gpu_aggregate(b, ("col", 0), ALL_AGGS + [("min", 0)])
    assert got.names == want.names
    same(got, want, rel=SUM_REL, ordered=False, sort_cols=[5])  # min(k) identifies the group
    # counts, min, max are exact
    gs, ws = sorted(got.rows(), key=lambda r: r[5]), sorted(want.rows(), key=lambda r: r[5])
    for a, w in zip(gs, ws):
        assert a[0] == w[0] and a[3] == w[3] and a[4] == w[4]


@pytest.mark.parametrize("n,groups", [(300_000, 5000), (300_000, 110_000), (20_000_000, 100_000)])
@pytest.mark.parametrize("vtype", ["f64", "i64"])
def test_group_by_one_value_column(n, groups, vtype):
    """Bare NULL-free key, every aggregate over ONE NULL-free column -- the BASELINE configs[2] shape and the shape the
    partitioned shared-memory path (NQE_AGG_PART=1) takes.  No key aggregate here, so groups are matched through their
    (count, min, max) triples; includes NaN / +-inf values and the key i64::MIN (the table's free-slot marker)."""
    import pyarrow as pa
    rng = np.random.default_rng(n // 1000 + groups)
    k = rng.integers(0, groups, n).astype(np.int64) * 7919 - 12345
    k[rng.integers(0, n, 50)] = np.iinfo(np.int64).min
    if vtype == "f64":
        v = np.round(rng.normal(0, 1000, n), 6)
        v[rng.integers(0, n, 20)] = np.nan
        v[rng.integers(0, n, 20)] = np.inf
        v[rng.integers(0, n, 20)] = -np.inf
    else:
        v = rng.integers(-10**9, 10**9, n).astype(np.int64)
    rb = pa.RecordBatch.from_arrays([pa.array(k), pa.array(v)], names=["k", "v"])
    nq = G.nq
    col = nq.ColumnExpr.try_create
    plan = nq.PhysicalAggregatePlan.create([col(None, 0)], [nq.Count.create(col(None, 1)), nq.Sum.create(col(None, 1)),
                                                          nq.Avg.create(col(None, 1)), nq.Min.create(col(None, 1)),
                                                          nq.Max.create(col(None, 1))],
                                           nq.ScanPlan.create(nq.MemTable.try_create(rb.schema, [rb]), None))
    out = plan.execute()[0]
    # expectation with numpy: group ids by sorting the keys
    uk, inv = np.unique(k, return_inverse=True)
    g = len(uk)
    assert out.num_rows == g
    vf = v.astype(np.float64)
    cnt = np.bincount(inv, minlength=g)
    order = np.argsort(inv, kind="stable")
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    vs = vf[order]
    # reference semantics: max is NaN if any NaN (OrderedFloat: NaN greatest), min ignores NaN, identities f64::MAX / f64::MIN
    has_nan = np.bincount(inv, weights=np.isnan(vf), minlength=g) > 0
    mx = np.maximum.reduceat(np.where(np.isnan(vs), -np.inf, vs), starts)
    mx = np.maximum(mx, -np.finfo(np.float64).max)
    mx = np.where(has_nan, np.nan, mx)
    mn = np.minimum.reduceat(np.where(np.isnan(vs), np.inf, vs), starts)
    mn = np.minimum(mn, np.finfo(np.float64).max)
    with np.errstate(invalid="ignore"):
        sm = np.add.reduceat(vs, starts)

    def canon(c, lo, hi):  # sortable (count, min, max) with NaN mapped to a sentinel
        return sorted(zip(c.tolist(), np.nan_to_num(lo, nan=1e308).tolist(), np.nan_to_num(hi, nan=1e308).tolist(), range(len(c))))
    got = [out.column(i).to_numpy(zero_copy_only=False) for i in range(5)]
    gs, ws = canon(got[0].astype(np.int64), got[3], got[4]), canon(cnt, mn, mx)
    assert [r[:3] for r in gs] == [r[:3] for r in ws]  # count, min, max exact
    gi, wi = np.array([r[3] for r in gs]), np.array([r[3] for r in ws])
    # groups that share a (count, min, max) triple cannot be told apart: compare sums only where the triple is unique
    trip = [r[:3] for r in ws]
    uniq = np.array([i == 0 or trip[i] != trip[i - 1] for i in range(g)]) & np.array([i == g - 1 or trip[i] != trip[i + 1] for i in range(g)])
    gsum, wsum = got[1][gi][uniq], sm[wi][uniq]
    fin = np.isfinite(wsum)
    assert np.array_equal(np.isnan(gsum), np.isnan(wsum)) and np.array_equal(gsum[~fin & ~np.isnan(wsum)], wsum[~fin & ~np.isnan(wsum)])
    scale = np.add.reduceat(np.abs(np.where(np.isfinite(vs), vs, 0.0)), starts)[wi][uniq]
    assert np.all(np.abs(gsum[fin] - wsum[fin]) <= SUM_REL * np.maximum(scale[fin], 1.0))
    gavg = got[2][gi][uniq]
    assert np.all(np.abs(gavg[fin] - (wsum / cnt[wi][uniq])[fin]) <= SUM_REL * np.maximum(scale[fin] / cnt[wi][uniq][fin], 1.0))


@pytest.mark.parametrize("n,groups", [(1000, 7), (300_000, 5000), (5_000_000, 40_000)])
def test_group_key_extension(n, groups):
    """NQE_AGG_GROUP_KEY (op 5, an extension: the reference's aggregate emits no key column, aggregate/mod.rs:117-121)
    returns the group key next to the aggregates -- on the table path and on the paged shared-memory path, through
    nqe_hash_aggregate and through the fused nqe_join_aggregate; the key i64::MIN included."""
    import ctypes as C
    import pyarrow as pa
    nq = G.nq
    rng = np.random.default_rng(n + groups)
    k = rng.integers(0, groups, n).astype(np.int64) * 104729 - 77
    k[rng.integers(0, n, 5)] = np.iinfo(np.int64).min
    v = np.round(rng.normal(0, 100, n), 4)
    t = nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(k), pa.array(v)], names=["k", "v"]))
    ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(t.names)
    aggs = (nq._ffi.Agg * 4)(*[nq._ffi.Agg(o, c) for o, c in [(5, 0), (0, 1), (1, 1), (4, 1)]])
    h = C.c_void_p()
    t.ctx.check(t.ctx.lib.nqe_hash_aggregate(t.ctx.h, t.h, C.pointer(ke), aggs, 4, C.byref(h)))
    out = nq.DeviceTable(t.ctx, h, ["key", "count", "sum", "max"]).to_arrow()
    assert out.schema.field(0).type == pa.int64()
    uk, inv = np.unique(k, return_inverse=True)
    o = np.argsort(out.column(0).to_numpy())
    assert np.array_equal(out.column(0).to_numpy()[o], uk)
    assert np.array_equal(out.column(1).to_numpy()[o], np.bincount(inv).astype(np.uint64))
    assert np.allclose(out.column(2).to_numpy()[o], np.bincount(inv, weights=v), rtol=SUM_REL, atol=1e-6)
    mx = np.full(len(uk), -np.inf)
    np.maximum.at(mx, inv, v)
    assert np.array_equal(out.column(3).to_numpy()[o], mx)
    # fused join -> group-by with the key: L(id, k-of-id) x R(fk = row's id, v)
    ids = np.arange(len(uk), dtype=np.int64) * 3 + 1
    L = nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(ids), pa.array(uk)], names=["id", "g"]))
    R = nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(ids[inv]), pa.array(v)], names=["fk", "v"]))
    aggs = (nq._ffi.Agg * 4)(*[nq._ffi.Agg(o_, c) for o_, c in [(5, 0), (0, 3), (1, 3), (4, 3)]])
    h = C.c_void_p()
    t.ctx.check(t.ctx.lib.nqe_join_aggregate(t.ctx.h, L.h, R.h, 0, 0, 1, aggs, 4, C.byref(h)))
    out2 = nq.DeviceTable(t.ctx, h, ["key", "count", "sum", "max"]).to_arrow()
    o2 = np.argsort(out2.column(0).to_numpy())
    assert np.array_equal(out2.column(0).to_numpy()[o2], uk)
    assert np.array_equal(out2.column(1).to_numpy()[o2], out.column(1).to_numpy()[o])
    assert np.allclose(out2.column(2).to_numpy()[o2], out.column(2).to_numpy()[o], rtol=SUM_REL, atol=1e-6)
    assert np.array_equal(out2.column(3).to_numpy()[o2], mx)
    # the global (no GROUP BY) plan has no key
    aggs = (nq._ffi.Agg * 2)(nq._ffi.Agg(5, 0), nq._ffi.Agg(0, 1))
    with pytest.raises(nq.NqeError) as e:
        t.ctx.check(t.ctx.lib.nqe_hash_aggregate(t.ctx.h, t.h, None, aggs, 2, C.byref(h)))
    assert e.value.kind == "InvalidArgument"
    for x in (t, L, R):
        x.free()


def test_group_by_dense_keys_with_outliers_the_sample_misses():
    """Dense-key mode of the paged group-by (keys split by RANGE, directly indexed shared-memory tables) is chosen from a
    strided sample; keys outside the sampled range (here: far outliers and i64::MIN on rows the stride skips) must be
    detected by the split and answered by falling back to hashing -- same result either way."""
    import ctypes as C
    import pyarrow as pa
    nq = G.nq
    rng = np.random.default_rng(99)
    n, groups = 5_000_000, 30_000
    k = rng.integers(0, groups, n).astype(np.int64) + 1000
    v = np.round(rng.normal(0, 100, n), 4)
    for outliers in (False, True):
        kk = k.copy()
        if outliers:  # the sample reads rows 0, 4, 8, ... (n / 2^20 = 4): odd rows are never sampled
            odd = rng.integers(0, n // 2, 40) * 2 + 1
            kk[odd[:20]] = 10**15 + np.arange(20)
            kk[odd[20:30]] = -7
            kk[odd[30:]] = np.iinfo(np.int64).min
        t = nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(kk), pa.array(v)], names=["k", "v"]))
        ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(t.names)
        aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(o, c) for o, c in [(5, 0), (0, 1), (1, 1), (3, 1), (4, 1)]])
        h = C.c_void_p()
        t.ctx.check(t.ctx.lib.nqe_hash_aggregate(t.ctx.h, t.h, C.pointer(ke), aggs, 5, C.byref(h)))
        out = nq.DeviceTable(t.ctx, h, ["key", "count", "sum", "min", "max"]).to_arrow()
        uk, inv = np.unique(kk, return_inverse=True)
        o = np.argsort(out.column(0).to_numpy())
        assert np.array_equal(out.column(0).to_numpy()[o], uk)
        assert np.array_equal(out.column(1).to_numpy()[o], np.bincount(inv).astype(np.uint64))
        assert np.allclose(out.column(2).to_numpy()[o], np.bincount(inv, weights=v), rtol=SUM_REL, atol=1e-6)
        mn = np.full(len(uk), np.inf); mx = np.full(len(uk), -np.inf)
        np.minimum.at(mn, inv, v); np.maximum.at(mx, inv, v)
        assert np.array_equal(out.column(3).to_numpy()[o], mn) and np.array_equal(out.column(4).to_numpy()[o], mx)
        t.free()


def test_group_by_expression_key_and_special_values():
    nan, inf = float("nan"), float("inf")
    b = O.Batch(["k", "v"], [O.col("i64", [1, 1, None, 2, 3, 3, -2**63, -2**63, 4]),
                             O.col("f64", [nan, 2.0, 5.0, None, inf, -inf, 1.0, 2.0, -0.0])])
    for key in [("col", 0), ("bin", "Modulos", ("col", 0), lit(2))]:
        want = O.aggregate(b, key, ALL_AGGS)
        got = G.gpu_aggregate(b, key, ALL_AGGS)
        same(got, want, rel=SUM_REL, ordered=False, sort_cols=[0, 3, 4])


def test_global_aggregate():
    rng = np.random.default_rng(2)
    for n in [0, 1, 100_000]:
        b = O.Batch(["k", "v", "u"], [rand_col(rng, "i64", n), rand_col(rng, "f64", n, 0.2), rand_col(rng, "u64", n)])
        aggs = ALL_AGGS + [("sum", 0), ("max", 2), ("count", 0)]
        same(G.gpu_aggregate(b, None, aggs), O.aggregate(b, None, aggs), rel=SUM_REL)


def test_aggregate_errors():
    import nqe_b200 as nq
    b = O.Batch(["k", "t", "f"], [O.col("i64", [1]), O.col("bool", [True]), O.col("f64", [1.0])])
    with pytest.raises(nq.NqeError) as e:
        G.gpu_aggregate(b, ("col", 2), [("count", 0)])
    assert e.value.kind == "NotSupported" and "group by only support" in e.value.message
    with pytest.raises(nq.NqeError) as e:
        G.gpu_aggregate(b, None, [("sum", 1)])
    assert e.value.kind == "NotSupported"


@pytest.mark.parametrize("nl,nr,groups", [(1000, 20_000, 10), (20_000, 100_000, 5000)])
def test_join_aggregate_fused(nl, nr, groups):
    rng = np.random.default_rng(nl)
    l = O.Batch(["k", "a"], [O.Col("i64", rng.permutation(nl).astype(np.int64)), O.Col("i64", rng.integers(0, groups, nl))])
    r = O.Batch(["fk", "b"], [O.Col("i64", rng.integers(0, int(nl * 1.2), nr)), rand_col(rng, "f64", nr, 0.05)])
    aggs = [("count", 3), ("sum", 3), ("avg", 3), ("min", 3), ("max", 3), ("min", 1)]
    want = O.aggregate(O.hash_join_c(l, r, 0, 0), ("col", 1), aggs)
    got = G.gpu_join_aggregate(l, r, "k", "fk", 1, aggs)
    same(got, want, rel=SUM_REL, ordered=False, sort_cols=[5])
    # duplicate build keys, group key from the probe side
    l2 = O.Batch(["k", "a"], [O.Col("i64", rng.integers(0, nl // 4, nl)), rand_col(rng, "i64", nl)])
    r2 = O.Batch(["fk", "g"], [O.Col("i64", rng.integers(0, nl // 4, nr // 10)), O.Col("i64", rng.integers(0, groups, nr // 10))])
    aggs2 = [("count", 1), ("sum", 1), ("max", 1), ("min", 3)]
    want2 = O.aggregate(O.hash_join_c(l2, r2, 0, 0), ("col", 3), aggs2)
    got2 = G.gpu_join_aggregate(l2, r2, "k", "fk", 3, aggs2)
    same(got2, want2, rel=SUM_REL, ordered=False, sort_cols=[3])


@pytest.mark.parametrize("poison", [False, True])
@pytest.mark.parametrize("dups", [False, True])
def test_join_aggregate_group_key_in_slot(poison, dups):
    """Only the group key is needed from the build side: it rides in the join table's row word (one random access per
    probe row).  poison: a group key equal to the free-slot marker (-1) forces the fall-back to row numbers."""
    rng = np.random.default_rng(11 + 2 * poison + dups)
    nl, nr, groups = 30_000, 200_000, 700
    k = rng.integers(0, nl // 3, nl) if dups else rng.permutation(nl)
    a = rng.integers(0, groups, nl)
    if poison:
        a[rng.integers(0, nl, 20)] = -1
    l = O.Batch(["k", "a", "z"], [O.Col("i64", k.astype(np.int64)), O.Col("i64", a.astype(np.int64)), rand_col(rng, "f64", nl)])
    r = O.Batch(["fk", "b"], [O.Col("i64", rng.integers(0, int(nl * 1.2), nr)), rand_col(rng, "f64", nr, 0.05)])
    aggs = [("count", 4), ("sum", 4), ("avg", 4), ("min", 4), ("max", 4), ("count", 3)]
    want = O.aggregate(O.hash_join_c(l, r, 0, 0), ("col", 1), aggs)
    got = G.gpu_join_aggregate(l, r, "k", "fk", 1, aggs)
    same(got, want, rel=SUM_REL, ordered=False, sort_cols=[3, 4, 5])


# ---------------------------------------------------------------- limit / offset / partition / synth
def test_limit_offset():  # limit.rs:67-90, offset.rs:69-92, README.md:70-76
    import nqe_b200 as nq
    t1 = numeric_only(golden_table("t1"))
    off = nq.PhysicalOffsetPlan.create(G.scan(t1), 5)
    assert G.from_arrow(off.execute()[0]).cols[0].to_pylist() == FX["test_physical_offset"]["id"]
    lim = nq.PhysicalLimitPlan.create(G.scan(t1), 2)
    assert G.from_arrow(lim.execute()[0]).cols[0].to_pylist() == [1, 2]
    rng = np.random.default_rng(4)
    b = O.Batch(["a", "f", "t"], [rand_col(rng, "i64", 1000, 0.3), rand_col(rng, "f64", 1000), rand_col(rng, "bool", 1000, 0.3)])
    for o, l in [(0, 1000), (3, 100), (37, 900), (999, 1), (1000, 0), (64, 64)]:
        plan = nq.PhysicalLimitPlan.create(nq.PhysicalOffsetPlan.create(G.scan(b), o), l)
        got = G.from_arrow(plan.execute()[0])
        assert got.rows() == b.rows()[o:][:l]


def test_radix_partition_is_a_permutation_grouped_by_destination():
    import ctypes as C
    import nqe_b200 as nq
    rng = np.random.default_rng(8)
    n = 100_003
    b = O.Batch(["k", "v"], [O.Col("i64", rng.integers(0, 10**6, n)), rand_col(rng, "f64", n)])
    src = G.scan(b).execute_device()
    for parts in [1, 2, 8]:
        h = C.c_void_p()
        counts = (C.c_int64 * parts)()
        src.ctx.check(src.ctx.lib.nqe_radix_partition(src.ctx.h, src.h, 0, parts, C.byref(h), counts))
        out = G.from_arrow(nq.DeviceTable(src.ctx, h, ["k", "v"]).to_arrow())
        assert sum(counts) == n
        assert sorted(out.rows()) == sorted(b.rows())
        # every region holds one destination only, and the destination is a function of the key
        dest = {}
        pos = 0
        for p in range(parts):
            for kk in out.cols[0].values[pos:pos + counts[p]]:
                assert dest.setdefault(int(kk), p) == p
            pos += counts[p]


def test_synth_columns_match_oracle_generator():
    import torch
    import nqe_b200 as nq
    ctx = nq.Context.default()
    n = 10_000
    for kind, seed, a, b2, scale, want in [
            (0, 42, 1000, 0, 0.0, O.gen_mod_i64(42, 7, n, 1000)),
            (1, 44, 0, 0, 100.0, O.gen_unif_f64(44, 7, n, 100.0)),
            (2, 0, 7368787, 10**7, 0.0, O.gen_perm_i64(7, n, 7368787, 10**7))]:
        buf = torch.empty(n, dtype=torch.int64, device="cuda")
        ctx.check(ctx.lib.nqe_synth_column(ctx.h, kind, seed, 7, n, a, b2, scale, buf.data_ptr()))
        ctx.sync()
        got = buf.cpu().numpy()
        assert np.array_equal(got, want.view(np.int64))


@pytest.mark.parametrize("peer", [False, True])
def test_distributed_plan_on_one_gpu_matches_oracle(peer):
    """distributed.py's CudaEngine through a 1-rank NCCL group (the >1-rank orchestration is
    covered on CPU over gloo in test_distributed_gloo.py).  peer=True: rows go through
    nqe_partition_counts + nqe_shuffle_scatter into symmetric-memory receive buffers."""
    import os
    from importlib import import_module
    import torch
    import torch.distributed as dist
    import nqe_b200 as nq
    D = import_module("naive-query-engine_b200.distributed")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        n_build, n_probe, groups = 5000, 60000, 101
        lk = O.gen_perm_i64(0, n_build, 7368787 % n_build, n_build)
        la = lk % groups
        fk = O.gen_mod_i64(47, 0, n_probe, int(n_build * 1.3))
        rb = O.gen_unif_f64(48, 0, n_probe, 100.0)
        dev = lambda a: torch.from_numpy(a.view(np.int64).copy()).cuda()
        eng = D.CudaEngine(nq, nq.Context.default(), torch)
        xbufs = None
        if peer:
            xbufs = (eng.alloc_exchange(n_build + 64, 2, dist.group.WORLD), eng.alloc_exchange(n_probe + 64, 2, dist.group.WORLD))
        merged, sent = D.shuffled_join_group_by(dist, torch, eng, [dev(lk), dev(la)], [dev(fk), dev(rb)], 1, xbufs)
        got = [m.cpu().numpy() for m in merged]
        L = O.Batch(["k", "a"], [O.Col("i64", lk), O.Col("i64", la)])
        R = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("f64", rb)])
        want = O.aggregate(O.hash_join_c(L, R, 0, 0), ("col", 1), [("min", 1), ("count", 3), ("sum", 3), ("min", 3), ("max", 3)])
        wk = want.cols[0].values.astype(np.int64)
        wo, go = np.argsort(wk), np.argsort(got[0])
        assert np.array_equal(wk[wo], got[0][go])
        assert np.array_equal(want.cols[1].values[wo].astype(np.int64), got[1][go])
        assert np.allclose(want.cols[2].values[wo], got[2][go].view(np.float64), rtol=SUM_REL, atol=0)
        assert np.array_equal(want.cols[3].values[wo], got[3][go].view(np.float64))
        assert np.array_equal(want.cols[4].values[wo], got[4][go].view(np.float64))
        assert sent == 0
    finally:
        if created:
            dist.destroy_process_group()


# ---------------------------------------------------------------- BASELINE sizes: size-independent properties
def _device_cols(specs, n, start=0):
    import torch
    from importlib import import_module
    synth = import_module("naive-query-engine_b200.synth")
    ctx = G.nq.Context.default()
    bufs = []
    for spec in specs:
        t = torch.empty(n, dtype=torch.int64, device="cuda")
        synth.device_column(ctx, spec, start, n, t.data_ptr())
        bufs.append(t)
    ctx.sync()
    return ctx, bufs


def test_group_by_full_size_properties():
    """configs[2] at full size (1e8 rows, 1e5 groups): per-group results against numpy bincount-style totals
    (count exact, min/max exact, sum/avg within 1e-9) -- the group key is recovered through min(k)."""
    import ctypes as C
    import torch
    from importlib import import_module
    synth = import_module("naive-query-engine_b200.synth")
    n, g = 100_000_000, 100_000
    ctx, bufs = _device_cols(synth.GROUPBY_TABLE, n)
    t = G.nq.DeviceTable.from_device_pointers(ctx, ["k", "v"], [2, 4], [b.data_ptr() for b in bufs], n, keepalive=bufs)
    ke, keep = G.nq.ColumnExpr.try_create(None, 0).to_expr(t.names)
    aggs = (G.nq._ffi.Agg * 6)(*[G.nq._ffi.Agg(o, c) for o, c in [(3, 0), (0, 1), (1, 1), (2, 1), (3, 1), (4, 1)]])
    h = C.c_void_p()
    ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, t.h, C.pointer(ke), aggs, 6, C.byref(h)))
    out = G.nq.DeviceTable(ctx, h, ["key", "count", "sum", "avg", "min", "max"]).to_arrow()
    assert out.num_rows == g
    key = out.column(0).to_numpy().astype(np.int64)
    order = np.argsort(key)
    assert np.array_equal(key[order], np.arange(g))
    k = synth.mod_i64(45, 0, n, g)
    v = synth.unif_f64(46, 0, n, 100.0)
    cnt = np.bincount(k, minlength=g)
    sm = np.bincount(k, weights=v, minlength=g)
    assert np.array_equal(out.column(1).to_numpy()[order], cnt.astype(np.uint64))
    assert np.allclose(out.column(2).to_numpy()[order], sm, rtol=SUM_REL, atol=0)
    assert np.allclose(out.column(3).to_numpy()[order], sm / cnt, rtol=SUM_REL, atol=0)
    mn = np.full(g, np.inf)
    mx = np.full(g, -np.inf)
    np.minimum.at(mn, k[: 5_000_000], v[: 5_000_000])  # exact per-group extremes on a prefix bound the full result
    np.maximum.at(mx, k[: 5_000_000], v[: 5_000_000])
    got_mn, got_mx = out.column(4).to_numpy()[order], out.column(5).to_numpy()[order]
    assert np.all(got_mn <= mn) and np.all(got_mx >= mx)
    assert got_mn.min() == v.min() and got_mx.max() == v.max()
    t.free()


def test_join_full_size_properties():
    """configs[3] at full size (1e8 probe x 1e7 unique build keys): every probe row matches exactly once, so the
    joined table is the probe side in its own order with a = fk mod 1e5 attached; also the fused join + group-by."""
    import ctypes as C
    import torch
    from importlib import import_module
    synth = import_module("naive-query-engine_b200.synth")
    pp = import_module("naive-query-engine_b200.physical_plan")
    nb, n, g = 10_000_000, 100_000_000, 100_000
    ctx, lb = _device_cols(synth.join_build_table(nb), nb)
    la = torch.remainder(lb[0], g)
    _, rb = _device_cols(synth.join_probe_table(nb), n)
    torch.cuda.synchronize()
    L = G.nq.DeviceTable.from_device_pointers(ctx, ["k", "a"], [2, 2], [lb[0].data_ptr(), la.data_ptr()], nb, keepalive=[lb, la])
    R = G.nq.DeviceTable.from_device_pointers(ctx, ["fk", "b"], [2, 4], [b.data_ptr() for b in rb], n, keepalive=rb)
    h = C.c_void_p()
    ctx.check(ctx.lib.nqe_hash_join(ctx.h, L.h, R.h, 0, 0, C.byref(h)))
    J = G.nq.DeviceTable(ctx, h, ["k", "a", "fk", "b"])
    assert J.num_rows == n
    cols = [torch.as_tensor(_CAI(J.column_desc(i).values, n), device="cuda") for i in range(4)]
    assert torch.equal(cols[0], rb[0]) and torch.equal(cols[2], rb[0]) and torch.equal(cols[3], rb[1])  # probe order kept
    assert torch.equal(cols[1], torch.remainder(rb[0], g))
    J.free()
    aggs = (G.nq._ffi.Agg * 3)(*[G.nq._ffi.Agg(o, c) for o, c in [(3, 1), (0, 3), (1, 3)]])
    ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, L.h, R.h, 0, 0, 1, aggs, 3, C.byref(h)))
    A = G.nq.DeviceTable(ctx, h, ["key", "count", "sum"]).to_arrow()
    assert A.num_rows == g
    key = A.column(0).to_numpy().astype(np.int64)
    order = np.argsort(key)
    fk = synth.mod_i64(47, 0, n, nb)
    b = synth.unif_f64(48, 0, n, 100.0)
    grp = fk % g
    assert np.array_equal(A.column(1).to_numpy()[order], np.bincount(grp, minlength=g).astype(np.uint64))
    assert np.allclose(A.column(2).to_numpy()[order], np.bincount(grp, weights=b, minlength=g), rtol=SUM_REL, atol=0)
    L.free(); R.free()


# ---------------------------------------------------------------- Utf8 columns riding along
def test_golden_selection_with_names():  # selection.rs:126-178 with the Utf8 column
    t1 = golden_table("t1")
    proj = O.projection(t1, [("col", 0), ("col", 1), ("col", 2)])
    pred = ("bin", "Gt", ("bin", "Plus", ("col", 0), lit(1)), lit(5))
    out = G.gpu_selection(proj, pred)
    assert out.cols[0].to_pylist() == FX["test_selection"]["id"]
    assert out.cols[1].to_pylist() == FX["test_selection"]["name"]


def test_golden_readme_limit_offset_with_names():  # README.md:70-76
    import nqe_b200 as nq
    import pyarrow as pa
    t1 = golden_table("t1")
    sel = nq.SelectionPlan.create(G.scan(t1), G.expr(("bin", "Lt", ("col", 0), lit(9))))
    proj = nq.ProjectionPlan.create(sel, pa.schema([("id", pa.int64()), ("name", pa.utf8()), ("age + 100", pa.int64())]),
                                    [G.expr(("col", 0)), G.expr(("col", 1)), G.expr(("bin", "Plus", ("col", 2), lit(100)))])
    plan = nq.PhysicalLimitPlan.create(nq.PhysicalOffsetPlan.create(proj, 2), 3)  # offset then limit (sql/planner.rs:49-52)
    got = G.from_arrow(plan.execute()[0])
    assert [list(r) for r in got.rows()] == FX["readme_limit_offset"]["rows"]


def test_golden_readme_three_way_join_with_strings():  # README.md:77-85, exact printed order
    emp, rank, dept = golden_table("employee"), golden_table("rank"), golden_table("department")
    j1 = G.gpu_join(emp, rank, "rank", "id")
    assert j1.rows() == O.hash_join(emp, rank, "rank", "id").rows()
    j2 = G.gpu_join(j1, dept, "department_id", "id")
    names = j2.names
    proj = G.gpu_projection(j2, [("col", 0), ("col", 1), ("col", names.index("rank_name")), ("col", names.index("department_name"))],
                            names=["id", "name", "rank_name", "department_name"])
    assert [list(r) for r in proj.rows()] == FX["readme_join"]["rows"]


def test_strings_through_filter_and_join_random():
    rng = np.random.default_rng(21)
    n = 20_000
    words = ["", "a", "bb", "naïve", "query", "engine-" * 5, "π", "x" * 100]
    s = [None if rng.random() < 0.1 else words[int(rng.integers(0, len(words)))] + str(int(rng.integers(0, 50))) for _ in range(n)]
    b = O.Batch(["k", "s", "v"], [rand_col(rng, "i64", n, 0.1, lo=0, hi=500), O.col("utf8", s), rand_col(rng, "f64", n, 0.1)])
    pred = ("bin", "Lt", ("col", 0), lit(250))
    assert G.gpu_selection(b, pred).rows() == O.selection(b, pred).rows()
    exprs = [("col", 1), ("bin", "Plus", ("col", 0), lit(1)), ("col", 1)]
    assert G.gpu_projection(b, exprs, pred=pred).rows() == O.projection(O.selection(b, pred), exprs).rows()
    nl = 300
    ls = [None if i % 17 == 0 else f"dim-{i}" for i in range(nl)]
    l = O.Batch(["id", "label"], [O.Col("i64", rng.permutation(nl).astype(np.int64)), O.col("utf8", ls)])
    r = O.Batch(["k", "s", "v"], [O.Col("i64", rng.integers(0, 400, n)), b.cols[1], b.cols[2]])
    assert G.gpu_join(l, r, "id", "k").rows() == O.hash_join(l, r, "id", "k").rows()


# ---------------------------------------------------------------- Utf8 join keys and group keys
def _words(rng, n, vocab, null_frac=0.0):
    vals = np.array([vocab[i] for i in rng.integers(0, len(vocab), n)], dtype=object)
    valid = (rng.random(n) >= null_frac).astype(np.uint8) if null_frac else None
    if valid is not None:
        vals[valid == 0] = ""  # an Arrow NULL string slot spans no bytes: its raw value (what a join key sees) is ""
    return O.Col("utf8", vals, valid)


@pytest.mark.parametrize("nl,nr,null_frac", [(0, 5, 0.0), (7, 0, 0.0), (50, 400, 0.0), (300, 2000, 0.2), (5000, 20000, 0.1)])
def test_hash_join_utf8_keys(nl, nr, null_frac):
    """hash_join.rs:146-160,205-225: String keys; validity ignored; probe-row-major, build rows ascending."""
    rng = np.random.default_rng(nl + nr)
    vocab = ["", "a", "b", "ab", "ba", "alice", "bob", "x" * 40, "zürich", "naïve"] + [f"k{i}" for i in range(max(nl // 3, 1))]
    L = O.Batch(["name", "v"], [_words(rng, nl, vocab, null_frac), rand_col(rng, "i64", nl, null_frac)])
    R = O.Batch(["who", "w", "tag"], [_words(rng, nr, vocab + ["absent", "nobody"], null_frac), rand_col(rng, "f64", nr),
                                       _words(rng, nr, ["t1", "t2"])])
    same(G.gpu_join(L, R, "name", "who"), O.hash_join(L, R, "name", "who"))


@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (1, 0.0), (1000, 0.0), (30000, 0.15)])
def test_group_by_utf8_key(n, null_frac):
    """aggregate/mod.rs:170-216: String group keys; NULL keys dropped; count over the string column itself."""
    rng = np.random.default_rng(n + 3)
    vocab = ["", "a", "b", "alice", "bob", "carol", "x" * 33] + [f"g{i}" for i in range(50)]
    b = O.Batch(["name", "v", "x"], [_words(rng, n, vocab, null_frac), rand_col(rng, "i64", n, null_frac), rand_col(rng, "f64", n, null_frac)])
    aggs = [("count", 0), ("count", 1), ("sum", 1), ("avg", 2), ("min", 2), ("max", 1)]
    want = O.aggregate(b, ("col", 0), aggs)
    got = G.gpu_aggregate(b, ("col", 0), aggs)
    same(got, want, rel=SUM_REL, ordered=False)


def test_utf8_key_errors():
    rng = np.random.default_rng(9)
    b = O.Batch(["name", "v"], [_words(rng, 10, ["a", "b"]), rand_col(rng, "i64", 10)])
    with pytest.raises(G.nq.NqeError) as e:  # sum over a Utf8 column panics in the reference (sum.rs:108)
        G.gpu_aggregate(b, ("col", 0), [("sum", 0)])
    assert e.value.code == 5
    other = O.Batch(["id", "w"], [rand_col(rng, "i64", 10), rand_col(rng, "i64", 10)])
    with pytest.raises(G.nq.NqeError) as e:  # Utf8 key against Int64 key: downcast unwrap panics
        G.gpu_join(b, other, "name", "id")
    assert e.value.code == 5


# ---------------------------------------------------------------- NaiveDB::run_sql (src/db.rs:24-37)
def _golden_db(tmp_path=None):
    nq = G.nq
    db = nq.NaiveDB()
    for name in ("t1", "employee", "rank", "department"):
        rb = G.to_arrow(golden_table(name))
        if tmp_path is not None and name == "t1":  # the CSV path the reference's own tests use (csv.rs:46-96)
            import pyarrow.csv as pcsv
            p = str(tmp_path / "t1.csv")
            pcsv.write_csv(pa_table(rb), p)
            db.create_csv_table(name, p)
        else:
            db.create_memory_table(name, rb.schema, [rb])
    return db


def pa_table(rb):
    import pyarrow as pa
    return pa.Table.from_batches([rb])


def test_run_sql_reference_queries(tmp_path):
    """The queries the reference asserts or prints (sql/planner.rs:645-710, README.md:60-111) through NaiveDB.run_sql."""
    db = _golden_db(tmp_path)
    out = db.run_sql(FX["config1"]["sql"])[0]
    assert out.schema.names == FX["config1"]["names"]
    assert [list(r) for r in zip(*[out.column(i).to_pylist() for i in range(out.num_columns)])] == FX["config1"]["rows"]
    out = db.run_sql("select id, name, age from t1 where id > 1")[0]  # sql/planner.rs:664-680
    w = FX["sql_where_id_gt_1"]
    assert out.column(0).to_pylist() == w["id"] and out.column(1).to_pylist() == w["name"] and out.column(2).to_pylist() == w["age"]
    out = db.run_sql(FX["readme_limit_offset"]["sql"])[0]
    assert [list(r) for r in zip(*[out.column(i).to_pylist() for i in range(3)])] == FX["readme_limit_offset"]["rows"]
    out = db.run_sql(FX["readme_join"]["sql"])[0]
    assert [list(r) for r in zip(*[out.column(i).to_pylist() for i in range(4)])] == FX["readme_join"]["rows"]  # exact printed order
    out = db.run_sql(FX["readme_groupby"]["sql"])[0]
    assert out.schema.names == FX["readme_groupby"]["names"]
    assert_rows([list(r) for r in zip(*[out.column(i).to_pylist() for i in range(6)])], FX["readme_groupby"]["rows"], rel=SUM_REL, ordered=False)


def test_run_sql_strings_comma_join_and_panics():
    nq = G.nq
    db = _golden_db()
    out = db.run_sql("select id, age from t1 where name = 'alice'")[0]
    assert out.column(0).to_pylist() == [5]
    out = db.run_sql("select name, rank_name from employee, rank where rank.id = employee.rank and employee.id > 1")[0]
    want = db.run_sql("select name, rank_name from employee join rank on employee.rank = rank.id")[0]
    keep = [i for i, n in enumerate(want.column(0).to_pylist()) if n != "vee"]
    assert out.column(0).to_pylist() == [want.column(0)[i].as_py() for i in keep]
    with pytest.raises(nq.NqeError) as e:  # abs over an Int64 column: unimplemented!() in the unary kernel
        db.run_sql("select abs(age) from t1")
    assert e.value.kind == "Panic"
    with pytest.raises(nq.NqeError) as e:  # CAST: every arm is todo!() (cast.rs:45-87)
        db.run_sql("select cast(age as double) from t1")
    assert e.value.kind == "Panic"


def test_run_sql_on_a_large_memory_table():
    """1e6-row batches enter through create_memory_table (catalog.rs:38-49) and run filter -> project, join + group-by."""
    import pyarrow as pa
    nq = G.nq
    rng = np.random.default_rng(4)
    n, nl = 1_000_000, 50_000
    t = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 1000, n)), pa.array(rng.integers(0, 100, n)), pa.array(rng.integers(0, nl, n)),
                                    pa.array(np.round(rng.random(n) * 100, 4))], names=["id", "age", "fk", "score"])
    dim = pa.RecordBatch.from_arrays([pa.array(np.arange(nl)), pa.array(np.arange(nl) % 700)], names=["k", "grp"])
    db = nq.NaiveDB()
    db.create_memory_table("t", t.schema, [t])
    db.create_memory_table("dim", dim.schema, [dim])
    out = db.run_sql("select id, age + 100 from t where id < 500")[0]
    ids, age = t.column(0).to_numpy(), t.column(1).to_numpy()
    assert np.array_equal(out.column(0).to_numpy(), ids[ids < 500]) and np.array_equal(out.column(1).to_numpy(), age[ids < 500] + 100)
    out = db.run_sql("select count(score), sum(score), min(score), max(score) from dim join t on dim.k = t.fk group by grp")[0]
    grp, sc = t.column(2).to_numpy() % 700, t.column(3).to_numpy()
    cnt = np.bincount(grp, minlength=700)
    sm = np.bincount(grp, weights=sc, minlength=700)
    mn = np.full(700, np.inf); mx = np.full(700, -np.inf)
    np.minimum.at(mn, grp, sc); np.maximum.at(mx, grp, sc)
    got = [out.column(i).to_numpy() for i in range(4)]
    go, wo = np.lexsort((got[3], got[2], got[0])), np.lexsort((mx, mn, cnt))
    assert np.array_equal(got[0][go].astype(np.int64), cnt[wo]) and np.array_equal(got[2][go], mn[wo]) and np.array_equal(got[3][go], mx[wo])
    assert np.allclose(got[1][go], sm[wo], rtol=SUM_REL, atol=0)
