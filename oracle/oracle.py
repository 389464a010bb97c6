"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

A restatement of naive-query-engine's physical_plan operators (reference
`src/physical_plan/*.rs`).  Numeric work is done by `libnqe_oracle.so`
(`nqe_oracle.c`, plain C, single thread); Utf8 columns -- which only ride along
through selection / take -- are handled here with plain Python loops.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module.  The product
package never does.

Expressions are nested tuples (a neutral wire form used by the tests for both
the oracle and the CUDA path):
    ("col", idx)
    ("lit", dtype, value_or_None)         dtype in {"bool","i64","u64","f64"}
    ("bin", op, left, right)              op in Operator names below
    ("un",  fn, child)                    fn in {"abs","sin","cos","tan"}
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnqe_oracle.so")

DT = {"bool": 1, "i64": 2, "u64": 3, "f64": 4}
DT_INV = {v: k for k, v in DT.items()}
NP = {"bool": np.uint8, "i64": np.int64, "u64": np.uint64, "f64": np.float64}
# reference src/logical_plan/expression.rs:335-362
OPS = ["Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq", "Plus", "Minus", "Multiply", "Divide",
       "Modulos", "And", "Or"]
OP_SYM = {"Eq": "=", "NotEq": "!=", "Lt": "<", "LtEq": "<=", "Gt": ">", "GtEq": ">=",
          "Plus": "+", "Minus": "-", "Multiply": "*", "Divide": "/", "Modulos": "%",
          "And": "and", "Or": "or"}
UNS = ["abs", "sin", "cos", "tan"]
AGGS = ["count", "sum", "avg", "min", "max"]

STATUS = {0: "OK", 1: "ArrowError(DivideByZero)", 2: "IntervalError", 3: "NotSupported",
          4: "NotImplemented", 5: "Panic"}


class OracleError(Exception):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.kind = STATUS.get(code, str(code))
        self.msg = msg


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "nqe_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _CCol(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("_pad", C.c_int32), ("len", C.c_int64),
                ("values", C.c_void_p), ("valid", C.c_void_p)]


class _CLit(C.Union):
    _fields_ = [("i", C.c_int64), ("u", C.c_uint64), ("f", C.c_double)]


class _CNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("col", C.c_int32), ("dtype", C.c_int32),
                ("is_null", C.c_int32), ("_pad", C.c_int32), ("lit", _CLit)]


class _CAgg(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.nqo_xxh64_u64.restype = C.c_uint64
        L.nqo_xxh64_u64.argtypes = [C.c_uint64]
        L.nqo_splitmix.restype = C.c_uint64
        L.nqo_splitmix.argtypes = [C.c_uint64, C.c_uint64]
        L.nqo_gen_mod_i64.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]
        L.nqo_gen_unif_f64.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_void_p]
        L.nqo_gen_perm_i64.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.nqo_free_col.argtypes = [C.POINTER(_CCol)]
        L.nqo_eval_expr.argtypes = [C.POINTER(_CCol), C.c_int, C.c_int64, C.POINTER(_CNode), C.c_int,
                                    C.POINTER(_CCol), C.c_char_p, C.c_int]
        L.nqo_selection.argtypes = [C.POINTER(_CCol), C.c_int, C.c_int64, C.POINTER(_CCol),
                                    C.POINTER(_CCol), C.POINTER(C.c_int64)]
        L.nqo_hash_join.argtypes = [C.POINTER(_CCol), C.c_int, C.c_int64, C.POINTER(_CCol), C.c_int,
                                    C.c_int64, C.c_int, C.c_int, C.POINTER(_CCol), C.POINTER(C.c_int64)]
        L.nqo_aggregate.argtypes = [C.POINTER(_CCol), C.c_int, C.c_int64, C.POINTER(_CCol),
                                    C.POINTER(_CAgg), C.c_int, C.POINTER(_CCol), C.POINTER(C.c_int64),
                                    C.c_char_p, C.c_int]
        _lib = L
    return _lib


# --------------------------------------------------------------------------
# data model
# --------------------------------------------------------------------------
@dataclass
class Col:
    """One Arrow-like column.  values: numpy array (bool -> uint8 0/1, utf8 ->
    object array of str); valid: uint8 array (1 = valid) or None (no nulls)."""
    dtype: str
    values: np.ndarray
    valid: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.values)

    def to_pylist(self):
        out = []
        for i in range(len(self.values)):
            if self.valid is not None and not self.valid[i]:
                out.append(None)
            elif self.dtype == "bool":
                out.append(bool(self.values[i]))
            elif self.dtype == "utf8":
                out.append(self.values[i])
            else:
                out.append(self.values[i].item())
        return out


def col(dtype: str, data, valid=None) -> Col:
    """Build a Col from a python list (None = NULL) or numpy array."""
    if isinstance(data, np.ndarray) and valid is None and dtype != "utf8":
        return Col(dtype, np.ascontiguousarray(data, dtype=NP[dtype]), None)
    data = list(data)
    if valid is None and any(v is None for v in data):
        valid = np.array([0 if v is None else 1 for v in data], dtype=np.uint8)
        fill = "" if dtype == "utf8" else 0
        data = [fill if v is None else v for v in data]
    if dtype == "utf8":
        arr = np.empty(len(data), dtype=object)
        arr[:] = data
        return Col(dtype, arr, None if valid is None else np.asarray(valid, dtype=np.uint8))
    return Col(dtype, np.array(data, dtype=NP[dtype]),
               None if valid is None else np.asarray(valid, dtype=np.uint8))


@dataclass
class Batch:
    names: list
    cols: list

    @property
    def num_rows(self):
        return len(self.cols[0]) if self.cols else 0

    def column(self, name_or_idx) -> Col:
        if isinstance(name_or_idx, int):
            return self.cols[name_or_idx]
        return self.cols[self.names.index(name_or_idx)]  # first match, schema.rs:116-131

    def rows(self):
        lists = [c.to_pylist() for c in self.cols]
        return [tuple(l[i] for l in lists) for i in range(self.num_rows)]


def _ccol(c: Col, keep: list) -> _CCol:
    v = np.ascontiguousarray(c.values)
    keep.append(v)
    s = _CCol(DT[c.dtype], 0, len(v), v.ctypes.data, None)
    if c.valid is not None:
        m = np.ascontiguousarray(c.valid, dtype=np.uint8)
        keep.append(m)
        s.valid = m.ctypes.data
    return s


def _take_out(cc: _CCol) -> Col:
    dt = DT_INV[cc.dtype]
    n = cc.len
    npdt = NP[dt]
    if n:
        buf = (C.c_uint8 * (n * np.dtype(npdt).itemsize)).from_address(cc.values)
        vals = np.frombuffer(buf, dtype=npdt).copy()
    else:
        vals = np.empty(0, dtype=npdt)
    valid = None
    if cc.valid:
        valid = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(cc.valid), dtype=np.uint8)[:n].copy()
    lib().nqo_free_col(C.byref(cc))
    return Col(dt, vals, valid)


# --------------------------------------------------------------------------
# expressions
# --------------------------------------------------------------------------
def _flatten(expr, out: list):
    k = expr[0]
    if k == "col":
        out.append(_CNode(0, 0, int(expr[1]), 0, 0, 0, _CLit(i=0)))
    elif k == "lit":
        _, dt, v = expr
        lit = _CLit(i=0)
        if v is not None:
            if dt == "f64":
                lit.f = float(v)
            elif dt == "u64":
                lit.u = int(v)
            elif dt == "bool":
                lit.u = 1 if v else 0
            else:
                lit.i = int(v)
        out.append(_CNode(1, 0, 0, DT[dt], 1 if v is None else 0, 0, lit))
    elif k == "bin":
        _flatten(expr[2], out)
        _flatten(expr[3], out)
        out.append(_CNode(2, OPS.index(expr[1]), 0, 0, 0, 0, _CLit(i=0)))
    elif k == "un":
        _flatten(expr[2], out)
        out.append(_CNode(3, UNS.index(expr[1]), 0, 0, 0, 0, _CLit(i=0)))
    else:
        raise ValueError(expr)


def expr_name(expr, names: Sequence[str]) -> str:
    """Output field name, logical_plan/expression.rs:236-331 (`age + 100`)."""
    k = expr[0]
    if k == "col":
        return names[expr[1]]
    if k == "lit":
        v = expr[2]
        return "null" if v is None else str(v)
    if k == "bin":
        return f"{expr_name(expr[2], names)} {OP_SYM[expr[1]]} {expr_name(expr[3], names)}"
    return f"{expr[1]}({expr_name(expr[2], names)})"


_CMP = {"Eq": lambda c: c == 0, "NotEq": lambda c: c != 0, "Lt": lambda c: c < 0, "LtEq": lambda c: c <= 0,
        "Gt": lambda c: c > 0, "GtEq": lambda c: c >= 0}


def _utf8_leaf(e, batch: Batch) -> bool:
    return (e[0] == "col" and batch.cols[e[1]].dtype == "utf8") or (e[0] == "lit" and e[1] == "utf8")


def _utf8_compare(op: str, l, r, batch: Batch) -> Col:
    """eq_dyn / neq_dyn / lt_dyn / lt_eq_dyn / gt_dyn / gt_eq_dyn on Utf8 arrays (binary.rs:127-132; literals are
    materialised to n copies first, binary.rs:123-124): bytewise order, NULL where either side is NULL."""
    n = batch.num_rows

    def rows(e):
        if e[0] == "lit":
            return [e[2]] * n
        return batch.cols[e[1]].to_pylist()
    a, b = rows(l), rows(r)
    vals, valid = np.zeros(n, dtype=np.uint8), np.ones(n, dtype=np.uint8)
    for i in range(n):
        if a[i] is None or b[i] is None:
            valid[i] = 0
            continue
        x, y = a[i].encode("utf-8"), b[i].encode("utf-8")
        vals[i] = 1 if _CMP[op]((x > y) - (x < y)) else 0
    return Col("bool", vals, None if valid.all() else valid)


def _rewrite_utf8(expr, batch: Batch):
    """comparisons of Utf8 leaves -> Boolean columns appended to a copy of the batch (what the numeric C evaluator sees)"""
    if expr[0] == "bin":
        if _utf8_leaf(expr[2], batch) and _utf8_leaf(expr[3], batch):
            if expr[1] in _CMP:
                batch.cols.append(_utf8_compare(expr[1], expr[2], expr[3], batch))
                batch.names.append(f"__utf8_cmp{len(batch.cols)}")
                return ("col", len(batch.cols) - 1)
            if expr[1] in ("And", "Or"):  # binary.rs:30-44
                raise OracleError(2, f"Cannot evaluate binary expression {expr[1]} with types Utf8 and Utf8")
            raise OracleError(5, "not implemented: arithmetic on Utf8")  # binary.rs:46-88 unimplemented!()
        if _utf8_leaf(expr[2], batch) != _utf8_leaf(expr[3], batch) and (expr[2][0] in ("col", "lit")) and (expr[3][0] in ("col", "lit")):
            names = {"bool": "Boolean", "i64": "Int64", "u64": "UInt64", "f64": "Float64", "utf8": "Utf8"}
            dt = lambda e: names[batch.cols[e[1]].dtype if e[0] == "col" else e[1]]
            raise OracleError(2, f"Cannot evaluate binary expression {expr[1]} with types {dt(expr[2])} and {dt(expr[3])}")
        return ("bin", expr[1], _rewrite_utf8(expr[2], batch), _rewrite_utf8(expr[3], batch))
    if expr[0] == "un":
        return ("un", expr[1], _rewrite_utf8(expr[2], batch))
    return expr


def _has_utf8_leaf(expr, batch: Batch) -> bool:
    if expr[0] in ("col", "lit"):
        return _utf8_leaf(expr, batch)
    return any(_has_utf8_leaf(e, batch) for e in expr[2:])


def evaluate(expr, batch: Batch) -> Col:
    """PhysicalExpr::evaluate(..).into_array() on one batch."""
    if expr[0] == "col" and batch.cols[expr[1]].dtype == "utf8":
        c = batch.cols[expr[1]]
        return Col("utf8", c.values.copy(), None if c.valid is None else c.valid.copy())
    if expr[0] != "col" and _has_utf8_leaf(expr, batch):
        batch = Batch(list(batch.names), list(batch.cols))
        expr = _rewrite_utf8(expr, batch)
    nodes: list = []
    _flatten(expr, nodes)
    keep: list = []
    # utf8 columns cannot take part in numeric expressions here; give the C side
    # placeholders so column indices line up.
    ccols = (_CCol * max(len(batch.cols), 1))()
    for i, c in enumerate(batch.cols):
        if c.dtype == "utf8":
            ccols[i] = _CCol(0, 0, len(c), None, None)
        else:
            ccols[i] = _ccol(c, keep)
    prog = (_CNode * len(nodes))(*nodes)
    out = _CCol()
    err = C.create_string_buffer(256)
    rc = lib().nqo_eval_expr(ccols, len(batch.cols), batch.num_rows, prog, len(nodes), C.byref(out), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return _take_out(out)


# --------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------
def selection(batch: Batch, predicate) -> Batch:
    """SelectionPlan::execute, selection.rs:58-107 (single-batch input)."""
    mask = evaluate(predicate, batch)
    if mask.dtype != "bool":
        raise OracleError(5, "predicate is not Boolean (downcast_ref unwrap)")
    out_cols: list = [None] * len(batch.cols)
    num = [i for i, c in enumerate(batch.cols) if c.dtype != "utf8"]
    if num:
        keep: list = []
        cc = (_CCol * len(num))(*[_ccol(batch.cols[i], keep) for i in num])
        cm = _ccol(mask, keep)
        oc = (_CCol * len(num))()
        nrows = C.c_int64(0)
        rc = lib().nqo_selection(cc, len(num), batch.num_rows, C.byref(cm), oc, C.byref(nrows))
        if rc:
            raise OracleError(rc)
        for j, i in enumerate(num):
            out_cols[i] = _take_out(oc[j])
    for i, c in enumerate(batch.cols):
        if c.dtype != "utf8":
            continue
        vals, valid = [], []
        for r in range(batch.num_rows):  # selection.rs:82-97
            mv = mask.valid is None or mask.valid[r]
            if mv:
                if mask.values[r]:
                    ok = c.valid is None or c.valid[r]
                    vals.append(c.values[r] if ok else "")
                    valid.append(1 if ok else 0)
            else:
                vals.append("")
                valid.append(0)
        arr = np.empty(len(vals), dtype=object)
        arr[:] = vals
        out_cols[i] = Col("utf8", arr, None if all(valid) else np.array(valid, dtype=np.uint8))
    return Batch(list(batch.names), out_cols)


def projection(batch: Batch, exprs: Sequence, names: Optional[Sequence[str]] = None) -> Batch:
    """ProjectionPlan::execute, projection.rs:43-70."""
    cols = [evaluate(e, batch) for e in exprs]
    if names is None:
        names = [expr_name(e, batch.names) for e in exprs]
    return Batch(list(names), cols)


def _take_utf8(c: Col, idx: np.ndarray) -> Col:
    vals = np.empty(len(idx), dtype=object)
    vals[:] = [c.values[i] for i in idx]
    valid = None if c.valid is None else c.valid[idx]
    if valid is not None and valid.all():
        valid = None
    return Col("utf8", vals, valid)


def hash_join(left: Batch, right: Batch, left_key: str, right_key: str) -> Batch:
    """HashJoin::execute = build + probe, hash_join.rs:124-254.  Keys are looked
    up by NAME (first match) in the left / right batch (:134-136, :171-172)."""
    lk, rk = left.names.index(left_key), right.names.index(right_key)
    ldt, rdt = left.cols[lk].dtype, right.cols[rk].dtype
    if ldt == "utf8" or rdt == "utf8":
        if ldt != rdt:
            raise OracleError(5, "key dtype mismatch (downcast unwrap)")
        table: dict = {}
        for i, v in enumerate(left.cols[lk].values):  # validity ignored, :146-160
            table.setdefault(v, []).append(i)
        outer, inner = [], []
        for i, v in enumerate(right.cols[rk].values):
            for j in table.get(v, ()):
                outer.append(j)
                inner.append(i)
        outer = np.array(outer, dtype=np.int64)
        inner = np.array(inner, dtype=np.int64)
    else:
        if ldt not in ("i64", "u64") or rdt not in ("i64", "u64"):
            raise OracleError(4, "join key dtype")
        # index-only join through C on the key columns, then take here
        keep: list = []
        lrow = Col("i64", np.arange(left.num_rows, dtype=np.int64))
        rrow = Col("i64", np.arange(right.num_rows, dtype=np.int64))
        lc = (_CCol * 2)(_ccol(left.cols[lk], keep), _ccol(lrow, keep))
        rc_ = (_CCol * 2)(_ccol(right.cols[rk], keep), _ccol(rrow, keep))
        oc = (_CCol * 4)()
        n = C.c_int64(0)
        rc = lib().nqo_hash_join(lc, 2, left.num_rows, rc_, 2, right.num_rows, 0, 0, oc, C.byref(n))
        if rc:
            raise OracleError(rc)
        outs = [_take_out(oc[i]) for i in range(4)]
        outer, inner = outs[1].values, outs[3].values
    cols = []
    for c, idx in [(c, outer) for c in left.cols] + [(c, inner) for c in right.cols]:
        if c.dtype == "utf8":
            cols.append(_take_utf8(c, idx))
        else:
            valid = None if c.valid is None else c.valid[idx]
            if valid is not None and valid.all():
                valid = None
            cols.append(Col(c.dtype, c.values[idx], valid))
    return Batch(list(left.names) + list(right.names), cols)


def hash_join_c(left: Batch, right: Batch, lkey: int, rkey: int) -> Batch:
    """Full join (take included) inside the C oracle: numeric columns only.
    Used as the timed CPU baseline."""
    keep: list = []
    lc = (_CCol * len(left.cols))(*[_ccol(c, keep) for c in left.cols])
    rc_ = (_CCol * len(right.cols))(*[_ccol(c, keep) for c in right.cols])
    oc = (_CCol * (len(left.cols) + len(right.cols)))()
    n = C.c_int64(0)
    rc = lib().nqo_hash_join(lc, len(left.cols), left.num_rows, rc_, len(right.cols), right.num_rows,
                             lkey, rkey, oc, C.byref(n))
    if rc:
        raise OracleError(rc)
    return Batch(list(left.names) + list(right.names), [_take_out(oc[i]) for i in range(len(oc))])


def aggregate(batch: Batch, group_expr, aggs: Sequence) -> Batch:
    """PhysicalAggregatePlan::execute, aggregate/mod.rs:113-222.
    aggs: sequence of (op_name, column_index).  group_expr None => global.
    Output has NO key column; names `sum(col)` etc. (sum.rs:57-66)."""
    names = [f"{op}({batch.names[ci]})" for op, ci in aggs]
    key = None
    if group_expr is not None:
        key = evaluate(group_expr, batch)
    if key is not None and key.dtype == "utf8":
        groups: dict = {}
        for i, v in enumerate(key.values):  # NULL keys dropped, mod.rs:178-186
            if key.valid is not None and not key.valid[i]:
                continue
            groups.setdefault(v, []).append(i)
        outs = [[] for _ in aggs]
        for rows in groups.values():
            idx = np.array(rows, dtype=np.int64)
            sub = Batch(batch.names, [Col(c.dtype, c.values[idx], None if c.valid is None else c.valid[idx])
                                      for c in batch.cols])
            one = aggregate(sub, ("lit", "i64", 0), aggs)
            for a in range(len(aggs)):
                outs[a].append(one.cols[a].values[0])
        return Batch(names, [Col("u64" if op == "count" else "f64",
                                 np.array(o, dtype=np.uint64 if op == "count" else np.float64))
                             for (op, _), o in zip(aggs, outs)])
    keep: list = []
    ccols = (_CCol * max(len(batch.cols), 1))()
    for i, c in enumerate(batch.cols):
        ccols[i] = _CCol(0, 0, len(c), None, None) if c.dtype == "utf8" else _ccol(c, keep)
    for op, ci in aggs:
        if batch.cols[ci].dtype == "utf8" and op != "count":
            raise OracleError(3 if key is None else 5, f"{op} func for Utf8 is not supported")
    cagg = (_CAgg * len(aggs))(*[_CAgg(AGGS.index(op), ci) for op, ci in aggs])
    oc = (_CCol * len(aggs))()
    ng = C.c_int64(0)
    err = C.create_string_buffer(256)
    # count over a utf8 column only needs validity: substitute a bool column
    for op, ci in aggs:
        if batch.cols[ci].dtype == "utf8":
            c = batch.cols[ci]
            ccols[ci] = _ccol(Col("bool", np.zeros(len(c), dtype=np.uint8), c.valid), keep)
    ck = _ccol(key, keep) if key is not None else None
    rc = lib().nqo_aggregate(ccols, len(batch.cols), batch.num_rows,
                             C.byref(ck) if ck is not None else None, cagg, len(aggs), oc, C.byref(ng),
                             err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return Batch(names, [_take_out(oc[i]) for i in range(len(aggs))])


# --------------------------------------------------------------------------
# synthetic data
# --------------------------------------------------------------------------
def gen_mod_i64(seed: int, start: int, n: int, mod: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().nqo_gen_mod_i64(seed, start, n, mod, out.ctypes.data)
    return out


def gen_unif_f64(seed: int, start: int, n: int, scale: float = 100.0) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    lib().nqo_gen_unif_f64(seed, start, n, scale, out.ctypes.data)
    return out


def gen_perm_i64(start: int, n: int, mul: int, mod: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().nqo_gen_perm_i64(start, n, mul, mod, out.ctypes.data)
    return out


def read_csv(path: str) -> Batch:
    """CsvTable::try_create, datasource/csv.rs:46-96, for the tiny fixtures:
    header row; dtype inferred per column (Int64 -> Float64 -> Utf8)."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip() != ""]
    names = lines[0].split(",")
    rows = [ln.split(",") for ln in lines[1:]]
    cols = []
    for j in range(len(names)):
        raw = [r[j] for r in rows]
        try:
            cols.append(col("i64", [int(x) for x in raw]))
            continue
        except ValueError:
            pass
        try:
            cols.append(col("f64", [float(x) for x in raw]))
            continue
        except ValueError:
            pass
        cols.append(col("utf8", raw))
    return Batch(names, cols)
