"""ctypes binding of the C ABI in include/nqe.h (libnqe_b200.so).

This is the same surface a Rust `extern "C"` block would bind.
There is no fallback: if the shared library is missing or no CUDA device is
present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnqe_b200.so")

BOOL, INT64, UINT64, FLOAT64, UTF8 = 1, 2, 3, 4, 5

STATUS_NAMES = {
    0: "OK", 1: "ArrowError(DivideByZero)", 2: "IntervalError", 3: "NotSupported", 4: "NotImplemented",
    5: "Panic", 6: "PlanError", 7: "LogicalError", 8: "InvalidArgument", 9: "CudaError", 10: "OutOfMemory",
}


class NqeError(Exception):
    """Mirror of the reference's ErrorCode (src/error.rs:13-40)."""

    def __init__(self, code: int, message: str = ""):
        self.code = code
        self.kind = STATUS_NAMES.get(code, str(code))
        self.message = message
        super().__init__(f"{self.kind}: {message}")


class ColumnDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("reserved", C.c_int32), ("length", C.c_int64), ("null_count", C.c_int64),
                ("values", C.c_void_p), ("validity", C.c_void_p), ("data", C.c_void_p), ("data_bytes", C.c_int64)]


class _Value(C.Union):
    _fields_ = [("i64", C.c_int64), ("u64", C.c_uint64), ("f64", C.c_double)]


class ExprNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("column", C.c_int32), ("dtype", C.c_int32),
                ("is_null", C.c_int32), ("reserved", C.c_int32), ("value", _Value)]


class Expr(C.Structure):
    _fields_ = [("nodes", C.POINTER(ExprNode)), ("n_nodes", C.c_int32), ("reserved", C.c_int32)]


class Agg(C.Structure):
    _fields_ = [("op", C.c_int32), ("column", C.c_int32)]


# every symbol include/nqe.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "nqe_abi_version": (C.c_int32, []),
    "nqe_device_count": (C.c_int32, []),
    "nqe_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(_P)]),
    "nqe_ctx_destroy": (None, [_P]),
    "nqe_last_error": (C.c_char_p, [_P]),
    "nqe_ctx_set_stream": (C.c_int32, [_P, _P]),
    "nqe_ctx_sync": (C.c_int32, [_P]),
    "nqe_ctx_kernel_launches": (C.c_int64, [_P]),
    "nqe_ctx_last_op_ms": (C.c_double, [_P]),
    "nqe_table_upload": (C.c_int32, [_P, C.POINTER(ColumnDesc), C.c_int32, C.POINTER(_P)]),
    "nqe_table_from_device": (C.c_int32, [_P, C.POINTER(ColumnDesc), C.c_int32, C.POINTER(_P)]),
    "nqe_table_num_rows": (C.c_int64, [_P]),
    "nqe_table_num_columns": (C.c_int32, [_P]),
    "nqe_table_column": (C.c_int32, [_P, C.c_int32, C.POINTER(ColumnDesc)]),
    "nqe_table_download_column": (C.c_int32, [_P, _P, C.c_int32, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64]),
    "nqe_table_free": (None, [_P]),
    "nqe_table_slice": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "nqe_table_concat": (C.c_int32, [_P, C.POINTER(_P), C.c_int32, C.POINTER(_P)]),
    "nqe_filter_project": (C.c_int32, [_P, _P, C.POINTER(Expr), C.POINTER(Expr), C.c_int32, C.POINTER(_P)]),
    "nqe_filter_project_host": (C.c_int32, [_P, C.POINTER(ColumnDesc), C.c_int32, C.POINTER(Expr), C.POINTER(Expr), C.c_int32,
                                            C.POINTER(ColumnDesc), C.POINTER(C.c_int64)]),
    "nqe_hash_join": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "nqe_hash_aggregate": (C.c_int32, [_P, _P, C.POINTER(Expr), C.POINTER(Agg), C.c_int32, C.POINTER(_P)]),
    "nqe_join_aggregate": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Agg), C.c_int32,
                                       C.POINTER(_P)]),
    "nqe_multi_create": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(_P)]),
    "nqe_multi_destroy": (None, [_P]),
    "nqe_multi_size": (C.c_int32, [_P]),
    "nqe_multi_ctx": (_P, [_P, C.c_int32]),
    "nqe_multi_last_error": (C.c_char_p, [_P]),
    "nqe_multi_table_copy": (C.c_int32, [_P, _P, C.c_int32, C.POINTER(_P)]),
    "nqe_multi_join_aggregate": (C.c_int32, [_P, _P, C.POINTER(_P), C.c_int32, C.c_int32, C.c_int32, C.POINTER(Agg), C.c_int32,
                                             C.POINTER(_P)]),
    "nqe_multi_hash_aggregate": (C.c_int32, [_P, C.POINTER(_P), C.POINTER(Expr), C.POINTER(Agg), C.c_int32, C.POINTER(_P)]),
    "nqe_radix_partition": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "nqe_partition_counts": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "nqe_shuffle_scatter": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "nqe_synth_column": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64,
                                     C.c_double, _P]),
}

_lib = None


def load():
    """dlopen libnqe_b200.so and bind every declared symbol.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
