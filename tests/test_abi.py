"""CPU-side checks of the C ABI: the library loads and exports every symbol that
include/nqe.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nqe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nqe_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import nqe_b200 as nq
    lib = C.CDLL(nq.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nqe.h but not exported"
    from importlib import import_module
    ffi = import_module("naive-query-engine_b200._ffi")
    assert set(names) == set(ffi.SYMBOLS), "ctypes binding and header disagree"
    assert nq.load().nqe_abi_version() == 1


def test_no_cpu_fallback_without_cuda():
    import nqe_b200 as nq
    if nq.load().nqe_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(nq.NqeError) as e:
        nq.Context(0)
    assert e.value.kind == "CudaError"
    with pytest.raises(nq.NqeError) as e:  # the multi-GPU entry has no fallback either
        nq.MultiContext([0, 0])
    assert e.value.kind == "CudaError"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "naive-query-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle-free", ""), f"{f} mentions the oracle"


def test_host_expression_lowering_is_postfix():
    import nqe_b200 as nq
    e = nq.PhysicalBinaryExpr.create(
        nq.PhysicalBinaryExpr.create(nq.ColumnExpr.try_create("id", None), "Plus",
                                     nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(1))),
        "Gt", nq.PhysicalLiteralExpr.create(nq.ScalarValue.Int64(5)))
    ex, keep = e.to_expr(["x", "id"])
    kinds = [(keep[i].kind, keep[i].op, keep[i].column) for i in range(ex.n_nodes)]
    assert kinds == [(0, 0, 1), (1, 0, 0), (2, 6, 0), (1, 0, 0), (2, 4, 0)]
    assert keep[1].value.i64 == 1 and keep[3].value.i64 == 5
    with pytest.raises(nq.NqeError) as err:
        nq.ColumnExpr.try_create(None, None)
    assert err.value.kind == "LogicalError"
    with pytest.raises(nq.NqeError):
        nq.ColumnExpr.try_create("nope", None).resolve(["a"])
