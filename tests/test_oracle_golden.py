"""Pins the CPU oracle against every vector the reference's own tests assert
for the hot path and against the README's printed results."""
import math
import struct

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_rows, fixtures, golden_table

FX = fixtures()["cases"]


def lit(v):
    return ("lit", "i64", v)


def test_xxh64_matches_published_algorithm():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(1)
    vals = [0, 1, 2**63 - 1, 2**64 - 1, 42] + [int(x) for x in rng.integers(0, 2**63, 50)]
    for v in vals:
        want = xxhash.xxh64_intdigest(struct.pack("<Q", v), seed=0)
        assert O.lib().nqo_xxh64_u64(v) == want


def test_splitmix_generator():
    def mix(x):
        M = (1 << 64) - 1
        x = (x + 0x9E3779B97F4A7C15) & M
        x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M
        x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M
        x ^= x >> 31
        return x
    got = O.gen_mod_i64(42, 5, 16, 1000)
    assert list(got) == [mix(42 + 5 + i) % 1000 for i in range(16)]
    f = O.gen_unif_f64(44, 0, 8, 100.0)
    assert list(f) == [100.0 * ((mix(44 + i) >> 11) * 2.0 ** -53) for i in range(8)]
    p = O.gen_perm_i64(0, 10**4, 7368787, 10**4)
    assert len(set(p.tolist())) == 10**4


def test_projection_vector():  # projection.rs:88-121
    t1 = golden_table("t1")
    out = O.projection(t1, [("bin", "Plus", ("col", 0), lit(1)), ("col", 1)])
    assert out.cols[0].to_pylist() == FX["test_projection"]["id_plus_1"]
    assert out.cols[1].to_pylist() == FX["test_projection"]["name"]


def test_selection_vector():  # selection.rs:126-178
    t1 = golden_table("t1")
    proj = O.projection(t1, [("col", 0), ("col", 1), ("col", 2)])
    pred = ("bin", "Gt", ("bin", "Plus", ("col", 0), lit(1)), lit(5))
    out = O.selection(proj, pred)
    assert out.cols[0].to_pylist() == FX["test_selection"]["id"]
    assert out.cols[1].to_pylist() == FX["test_selection"]["name"]


def test_sql_where_vector():  # sql/planner.rs:664-680
    t1 = golden_table("t1")
    sel = O.selection(t1, ("bin", "Gt", ("col", 0), lit(1)))
    out = O.projection(sel, [("col", 0), ("col", 1), ("col", 2)])
    assert out.cols[0].to_pylist() == FX["sql_where_id_gt_1"]["id"]
    assert out.cols[1].to_pylist() == FX["sql_where_id_gt_1"]["name"]
    assert out.cols[2].to_pylist() == FX["sql_where_id_gt_1"]["age"]


def test_abs_sin_vectors():  # unary.rs:123-170, bit-exact vs glibc
    t1 = golden_table("t1")
    assert O.evaluate(("un", "abs", ("col", 3)), t1).to_pylist() == FX["test_abs_expression"]["score"]
    assert O.evaluate(("un", "sin", ("col", 3)), t1).to_pylist() == FX["test_sin_expression"]["score"]
    # Tan is wired to cos (unary.rs:96)
    assert O.evaluate(("un", "tan", ("col", 3)), t1).to_pylist() == [math.cos(x) for x in t1.cols[3].values]


def test_config1():  # BASELINE.json configs[0]
    t1 = golden_table("t1")
    sel = O.selection(t1, ("bin", "Lt", ("col", 0), lit(9)))
    out = O.projection(sel, [("col", 0), ("bin", "Plus", ("col", 2), lit(100))])
    assert out.names == FX["config1"]["names"]
    assert [list(r) for r in out.rows()] == FX["config1"]["rows"]


def test_readme_limit_offset():  # README.md:70-76
    t1 = golden_table("t1")
    sel = O.selection(t1, ("bin", "Lt", ("col", 0), lit(9)))
    out = O.projection(sel, [("col", 0), ("col", 1), ("bin", "Plus", ("col", 2), lit(100))])
    rows = out.rows()[2:][:3]  # offset 2 then limit 3 (sql/planner.rs:49-52)
    assert [list(r) for r in rows] == FX["readme_limit_offset"]["rows"]


def test_readme_three_way_join_order():  # README.md:77-85
    emp, rank, dept = golden_table("employee"), golden_table("rank"), golden_table("department")
    j1 = O.hash_join(emp, rank, "rank", "id")
    # after join 1 names are id,name,department_id,rank,id,rank_name; "id" resolves to the first
    j2 = O.hash_join(j1, dept, "department_id", "id")
    names = j2.names
    proj = O.projection(j2, [("col", 0), ("col", 1), ("col", names.index("rank_name")),
                             ("col", names.index("department_name"))])
    assert [list(r) for r in proj.rows()] == FX["readme_join"]["rows"]


def test_readme_groupby_bit_exact():  # README.md:105-111
    t1 = golden_table("t1")
    out = O.aggregate(t1, ("bin", "Modulos", ("col", 0), lit(3)),
                      [("count", 0), ("sum", 2), ("sum", 3), ("avg", 3), ("max", 3), ("min", 3)])
    assert out.names == FX["readme_groupby"]["names"]
    assert out.cols[0].dtype == "u64" and out.cols[1].dtype == "f64"
    assert_rows([list(r) for r in out.rows()], FX["readme_groupby"]["rows"], rel=0.0, ordered=False)


# ---- semantics the reference's tests do not pin (parity unpinned): the oracle's
# ---- reading of the source, locked so the CUDA path is compared to a fixed target
def test_null_mask_rows_are_kept_as_null_rows():  # selection.rs:46
    b = O.Batch(["a", "b"], [O.col("i64", [1, None, 3, 4]), O.col("f64", [1.5, 2.5, None, 4.5])])
    out = O.selection(b, ("bin", "Gt", ("col", 0), lit(2)))
    assert out.rows() == [(None, None), (3, None), (4, 4.5)]


def test_type_mismatch_is_interval_error():  # binary.rs:114-119
    b = O.Batch(["a"], [O.col("i64", [1, 2])])
    with pytest.raises(O.OracleError) as e:
        O.evaluate(("bin", "Lt", ("col", 0), ("lit", "f64", 9.5)), b)
    assert e.value.kind == "IntervalError"
    assert "Cannot evaluate binary expression Lt with types Int64 and Float64" in e.value.msg


def test_divide_by_zero_and_wrapping():
    b = O.Batch(["a", "b"], [O.col("i64", [2**63 - 1, 7, -7]), O.col("i64", [1, 0, 2])])
    assert O.evaluate(("bin", "Plus", ("col", 0), ("col", 1)), b).to_pylist()[0] == -2**63
    with pytest.raises(O.OracleError) as e:
        O.evaluate(("bin", "Divide", ("col", 0), ("col", 1)), b)
    assert e.value.kind == "ArrowError(DivideByZero)"
    b2 = O.Batch(["a", "b"], [O.col("i64", [7, -7]), O.col("i64", [2, 2])])
    assert O.evaluate(("bin", "Modulos", ("col", 0), ("col", 1)), b2).to_pylist() == [1, -1]
    assert O.evaluate(("bin", "Divide", ("col", 0), ("col", 1)), b2).to_pylist() == [3, -3]
    # a NULL divisor slot holding 0 is not an error
    b3 = O.Batch(["a", "b"], [O.col("i64", [7, 8]), O.col("i64", [2, None])])
    assert O.evaluate(("bin", "Divide", ("col", 0), ("col", 1)), b3).to_pylist() == [3, None]


def test_kleene_logic():
    T, F, N = True, False, None
    a = O.col("bool", [T, T, T, F, F, F, N, N, N])
    b = O.col("bool", [T, F, N, T, F, N, T, F, N])
    bt = O.Batch(["a", "b"], [a, b])
    assert O.evaluate(("bin", "And", ("col", 0), ("col", 1)), bt).to_pylist() == [T, F, N, F, F, F, N, F, N]
    assert O.evaluate(("bin", "Or", ("col", 0), ("col", 1)), bt).to_pylist() == [T, T, T, T, F, N, T, N, N]


def test_aggregate_quirks():
    b = O.Batch(["k", "v"], [O.col("i64", [1, 1, None, 2]), O.col("f64", [float("nan"), 2.0, 5.0, None])])
    out = O.aggregate(b, ("col", 0), [("count", 1), ("sum", 1), ("avg", 1), ("min", 1), ("max", 1)])
    rows = out.rows()
    assert rows[0][0] == 2 and math.isnan(rows[0][1]) and rows[0][3] == 2.0 and math.isnan(rows[0][4])
    # group 2 has only a NULL value: count 0, sum 0, avg NaN, min f64::MAX, max f64::MIN
    assert rows[1][0] == 0 and rows[1][1] == 0.0 and math.isnan(rows[1][2])
    assert rows[1][3] == 1.7976931348623157e308 and rows[1][4] == -1.7976931348623157e308
    g = O.aggregate(b, None, [("count", 0), ("sum", 0)])
    assert g.rows() == [(3, 4.0)]


def test_join_ignores_key_validity_and_orders_probe_major():
    l = O.Batch(["k", "a"], [O.col("i64", [5, 7, 5, 0], valid=[1, 1, 1, 0]), O.col("i64", [10, 11, 12, 13])])
    r = O.Batch(["fk", "b"], [O.col("i64", [7, 5, 0, 9]), O.col("f64", [0.5, 1.5, 2.5, 3.5])])
    out = O.hash_join(l, r, "k", "fk")
    assert out.rows() == [(7, 11, 7, 0.5), (5, 10, 5, 1.5), (5, 12, 5, 1.5), (None, 13, 0, 2.5)]
    out2 = O.hash_join_c(l, r, 0, 0)
    assert out2.rows() == out.rows()
