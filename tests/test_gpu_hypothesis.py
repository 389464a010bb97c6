"""Property tests (hypothesis): the CUDA path against the CPU oracle on random small batches --
NULLs, duplicates, empty inputs, all-pass / all-fail masks, random expression trees."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O
from tests import gpu_helpers as G
from tests.helpers import assert_rows

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
I64_VALUES = st.integers(-6, 6)
F64_VALUES = st.sampled_from([-2.5, -1.0, -0.5, 0.0, 0.25, 1.0, 1.5, 3.75, 1e9, -1e9])


@st.composite
def batches(draw, min_rows=0, max_rows=200):
    n = draw(st.integers(min_rows, max_rows))
    nulls = draw(st.booleans())

    def column(dtype, values):
        v = draw(st.lists(values, min_size=n, max_size=n))
        valid = None
        if nulls and n:
            valid = np.array(draw(st.lists(st.booleans(), min_size=n, max_size=n)), dtype=np.uint8)
        arr = np.array(v, dtype={"i64": np.int64, "f64": np.float64, "bool": np.uint8}[dtype])
        return O.Col(dtype, arr, valid)

    return O.Batch(["a", "b", "x", "y", "f"], [column("i64", I64_VALUES), column("i64", I64_VALUES), column("f64", F64_VALUES),
                                             column("f64", F64_VALUES), column("bool", st.integers(0, 1))])


TYPES = {0: "i64", 1: "i64", 2: "f64", 3: "f64"}


@st.composite
def arith(draw, dtype, depth=2):
    cols = [i for i, t in TYPES.items() if t == dtype]
    if depth == 0 or draw(st.booleans()):
        if draw(st.booleans()):
            return ("col", draw(st.sampled_from(cols)))
        return ("lit", dtype, draw(I64_VALUES if dtype == "i64" else F64_VALUES))
    op = draw(st.sampled_from(["Plus", "Minus", "Multiply"]))
    left = draw(arith(dtype, depth - 1))
    if left[0] == "lit":  # keep a column on the left so that the result dtype is well defined
        left = ("col", draw(st.sampled_from(cols)))
    return ("bin", op, left, draw(arith(dtype, depth - 1)))


@st.composite
def predicate(draw, depth=2):
    if depth == 0 or draw(st.integers(0, 2)) == 0:
        if draw(st.integers(0, 4)) == 0:
            return ("col", 4)
        dtype = draw(st.sampled_from(["i64", "f64"]))
        op = draw(st.sampled_from(["Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq"]))
        return ("bin", op, draw(arith(dtype, 1)), draw(arith(dtype, 1)))
    return ("bin", draw(st.sampled_from(["And", "Or"])), draw(predicate(depth - 1)), draw(predicate(depth - 1)))


def same(a, b, **kw):
    assert [c.dtype for c in a.cols] == [c.dtype for c in b.cols]
    assert_rows([list(r) for r in a.rows()], [list(r) for r in b.rows()], **kw)


@settings(**SETTINGS)
@given(batches(), predicate(), st.lists(st.one_of(arith("i64"), arith("f64")), min_size=1, max_size=3))
def test_filter_project_property(b, pred, exprs):
    exprs = [e if e[0] != "lit" else ("col", 0) for e in exprs]
    want = O.projection(O.selection(b, pred), exprs)
    same(G.gpu_projection(b, exprs, pred=pred), want)


@settings(**SETTINGS)
@given(batches(), st.integers(1, 4), st.lists(st.tuples(st.sampled_from(["count", "sum", "avg", "min", "max"]), st.integers(0, 3)),
                                             min_size=1, max_size=5))
def test_group_by_property(b, mod, aggs):
    key = ("bin", "Modulos", ("col", 0), ("lit", "i64", mod))
    aggs = list(aggs) + [("min", 0), ("max", 0)]  # tie the (unordered) groups down
    want = O.aggregate(b, key, aggs)
    got = G.gpu_aggregate(b, key, aggs)
    same(got, want, rel=1e-9, ordered=False, sort_cols=[len(aggs) - 2, len(aggs) - 1])


@settings(**SETTINGS)
@given(batches(max_rows=60), batches(max_rows=120))
def test_hash_join_property(l, r):
    l = O.Batch(["k", "p", "x1", "y1", "f1"], l.cols)
    want = O.hash_join(l, r, "k", "a")
    same(G.gpu_join(l, r, "k", "a"), want)
