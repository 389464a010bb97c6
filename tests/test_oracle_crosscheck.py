"""Independent cross-check of the oracle: on NULL-free random data, where pyarrow 24's kernels and the
reference's arrow-rs 13 kernels agree by construction (no NULL masks, no zero divisors, no NaN), the oracle's
selection/projection, inner hash join and group-by must equal pyarrow.compute / Table.join / Table.group_by.
This does not pin the reference's quirks (those are covered by the golden vectors and the source-cited tests in
test_oracle_golden.py); it guards the bulk arithmetic of the restatement against an unrelated implementation."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import oracle as O


def _tables(seed, n, nl):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, 1000, n).astype(np.int64)
    age = rng.integers(0, 100, n).astype(np.int64)
    score = np.round(rng.random(n) * 100, 6)
    lk = rng.permutation(nl).astype(np.int64)
    la = (lk % 37).astype(np.int64)
    fk = rng.integers(0, int(nl * 1.3), n).astype(np.int64)
    return ids, age, score, lk, la, fk


@pytest.mark.parametrize("seed,n", [(1, 1000), (2, 50_000)])
def test_selection_projection_against_pyarrow(seed, n):
    ids, age, score, *_ = _tables(seed, n, 10)
    b = O.Batch(["id", "age", "score"], [O.Col("i64", ids), O.Col("i64", age), O.Col("f64", score)])
    pred = ("bin", "And", ("bin", "Lt", ("col", 0), ("lit", "i64", 500)), ("bin", "GtEq", ("col", 2), ("lit", "f64", 12.5)))
    exprs = [("col", 0), ("bin", "Plus", ("col", 1), ("lit", "i64", 100)), ("bin", "Multiply", ("col", 2), ("lit", "f64", 0.5)),
             ("bin", "Modulos", ("col", 0), ("lit", "i64", 7)), ("bin", "Minus", ("col", 1), ("col", 0))]
    got = O.projection(O.selection(b, pred), exprs)
    t = pa.table({"id": ids, "age": age, "score": score})
    mask = pc.and_(pc.less(t["id"], 500), pc.greater_equal(t["score"], 12.5))
    f = t.filter(mask)
    want = [f["id"], pc.add(f["age"], 100), pc.multiply(f["score"], 0.5),
            pa.array(np.fmod(f["id"].to_numpy(), 7).astype(np.int64)),  # Rust % = truncated remainder = fmod on non-negatives
            pc.subtract(f["age"], f["id"])]
    assert got.num_rows == f.num_rows
    for c, w in zip(got.cols, want):
        w = w.combine_chunks() if isinstance(w, pa.ChunkedArray) else w
        assert np.array_equal(c.values, w.to_numpy()), "column mismatch"


@pytest.mark.parametrize("seed,n,nl", [(3, 2000, 300), (4, 40_000, 5000)])
def test_inner_join_against_pyarrow(seed, n, nl):
    _, _, score, lk, la, fk = _tables(seed, n, nl)
    l = O.Batch(["k", "a"], [O.Col("i64", lk), O.Col("i64", la)])
    r = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("f64", score)])
    got = O.hash_join_c(l, r, 0, 0)
    want = pa.table({"k": lk, "a": la}).join(pa.table({"fk": fk, "b": score}), keys="k", right_keys="fk", join_type="inner",
                                             coalesce_keys=False)
    assert got.num_rows == want.num_rows
    g = sorted(zip(*[c.values.tolist() for c in got.cols]))
    w = sorted(zip(*[want[name].to_pylist() for name in ["k", "a", "fk", "b"]]))
    assert g == w
    # and the reference's order: probe-row-major (hash_join.rs:80-103) -- pyarrow's order is unspecified, so check it directly
    assert np.array_equal(got.cols[2].values, fk[np.isin(fk, lk)])


@pytest.mark.parametrize("seed,n,groups", [(5, 3000, 17), (6, 60_000, 2500)])
def test_group_by_against_pyarrow(seed, n, groups):
    rng = np.random.default_rng(seed)
    k = rng.integers(-groups // 2, groups // 2, n).astype(np.int64)
    v = np.round(rng.normal(0, 100, n), 6)
    got = O.aggregate(O.Batch(["k", "v"], [O.Col("i64", k), O.Col("f64", v)]), ("col", 0),
                      [("count", 1), ("sum", 1), ("avg", 1), ("min", 1), ("max", 1), ("min", 0)])
    want = pa.table({"k": k, "v": v}).group_by("k").aggregate([("v", "count"), ("v", "sum"), ("v", "mean"), ("v", "min"), ("v", "max")])
    wd = {kk: row for kk, *row in zip(want["k"].to_pylist(), want["v_count"].to_pylist(), want["v_sum"].to_pylist(),
                                      want["v_mean"].to_pylist(), want["v_min"].to_pylist(), want["v_max"].to_pylist())}
    rows = list(got.rows())
    assert len(rows) == len(wd)
    for cnt, sm, avg, mn, mx, key in rows:
        w = wd[int(key)]  # min(k) comes back as f64 (max.rs:61-75): exact for |k| < 2^53
        assert cnt == w[0] and mn == w[3] and mx == w[4]
        assert abs(sm - w[1]) <= 1e-9 * max(1.0, abs(w[1])) + 1e-7 and abs(avg - w[2]) <= 1e-9 * max(1.0, abs(w[2])) + 1e-9
