"""Shared test helpers: golden fixtures -> oracle Batches, result comparison."""
import json
import math
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.json")


def fixtures():
    with open(GOLDEN) as f:
        return json.load(f)


def golden_table(name) -> O.Batch:
    """Same dtype inference as the reference's CSV reader on these fixtures."""
    t = fixtures()["tables"][name]
    cols = []
    for j in range(len(t["names"])):
        raw = [r[j] for r in t["rows"]]
        try:
            cols.append(O.col("i64", [int(x) for x in raw]))
            continue
        except ValueError:
            pass
        try:
            cols.append(O.col("f64", [float(x) for x in raw]))
            continue
        except ValueError:
            pass
        cols.append(O.col("utf8", raw))
    return O.Batch(list(t["names"]), cols)


def _key(v):
    if v is None:
        return (0, 0)
    if isinstance(v, float):
        if math.isnan(v):
            return (2, 0)
        return (1, v)
    if isinstance(v, str):
        return (3, v)
    return (1, v)


def sort_rows(rows, cols=None):
    """Sort rows for multiset comparison.  `cols`: indices of columns that are
    exact (ints, counts, min/max, a key) -- float sums must not drive the order."""
    if cols is None:
        return sorted(rows, key=lambda r: tuple(_key(v) for v in r))
    return sorted(rows, key=lambda r: tuple(_key(r[c]) for c in cols))


def rows_close(a, b, rel=0.0):
    """Row-by-row comparison; ints/strings/None exact, floats within rel."""
    if len(a) != len(b):
        return False, f"row count {len(a)} != {len(b)}"
    for i, (ra, rb) in enumerate(zip(a, b)):
        if len(ra) != len(rb):
            return False, f"row {i} width"
        for x, y in zip(ra, rb):
            if isinstance(x, float) or isinstance(y, float):
                if x is None or y is None:
                    if x is not y:
                        return False, f"row {i}: {ra} vs {rb}"
                    continue
                if math.isnan(x) and math.isnan(y):
                    continue
                if x == y:
                    continue
                if rel and abs(x - y) <= rel * max(abs(x), abs(y)):
                    continue
                return False, f"row {i}: {ra} vs {rb}"
            elif x != y:
                return False, f"row {i}: {ra} vs {rb}"
    return True, ""


def assert_rows(a, b, rel=0.0, ordered=True, sort_cols=None):
    if not ordered:
        a, b = sort_rows(a, sort_cols), sort_rows(b, sort_cols)
    ok, why = rows_close(a, b, rel)
    assert ok, why
