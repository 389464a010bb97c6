"""Device-side tables: Arrow RecordBatch <-> HBM through the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import pyarrow as pa

from . import _ffi
from ._ffi import BOOL, FLOAT64, INT64, UINT64, UTF8, ColumnDesc, NqeError

_PA_TO_NQE = {pa.bool_(): BOOL, pa.int64(): INT64, pa.uint64(): UINT64, pa.float64(): FLOAT64, pa.utf8(): UTF8}
_NQE_TO_PA = {v: k for k, v in _PA_TO_NQE.items()}


def nqe_dtype(t: pa.DataType) -> int:
    if t == pa.large_utf8():
        raise NqeError(4, "LargeUtf8 is not an arrow type the reference produces")
    try:
        return _PA_TO_NQE[t]
    except KeyError:
        # reference: `_ => unimplemented!()` in selection.rs:98 / binary.rs:86
        raise NqeError(5, f"not implemented: data type {t}")


class Context:
    """One nqe_ctx: one GPU, one stream.  `Context.default()` is created lazily."""

    _default: Optional["Context"] = None

    def __init__(self, device: int = 0):
        self.lib = _ffi.load()
        h = C.c_void_p()
        rc = self.lib.nqe_ctx_create(device, C.byref(h))
        if rc != 0:
            raise NqeError(rc, "nqe_ctx_create failed: no usable CUDA device (the CUDA path is the only path)")
        self.h = h
        self.device = device

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context(0)
        return cls._default

    @classmethod
    def set_default(cls, ctx: "Context") -> None:
        """The context host-side plans (`plan.execute()` over MemTables) run on: one process per GPU sets its own."""
        cls._default = ctx

    def check(self, rc: int):
        if rc != 0:
            raise NqeError(rc, self.lib.nqe_last_error(self.h).decode())

    def set_stream(self, cuda_stream: int):
        self.check(self.lib.nqe_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self.check(self.lib.nqe_ctx_sync(self.h))

    @property
    def kernel_launches(self) -> int:
        return self.lib.nqe_ctx_kernel_launches(self.h)

    @property
    def last_op_ms(self) -> float:
        return self.lib.nqe_ctx_last_op_ms(self.h)

    def close(self):
        if self.h:
            self.lib.nqe_ctx_destroy(self.h)
            self.h = None


class MultiContext:
    """nqe_multi: several GPUs of one node driven from this one process (include/nqe.h).  `members[i]` is a Context
    borrowed from the nqe_multi (do not close it); tables for member i are created through `members[i]`."""

    def __init__(self, devices: Sequence[int]):
        self.lib = _ffi.load()
        arr = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.nqe_multi_create(arr, len(devices), C.byref(h))
        if rc != 0:
            raise NqeError(rc, "nqe_multi_create failed: no usable CUDA device (the CUDA path is the only path)")
        self.h = h
        self.members: List[Context] = []
        for i in range(len(devices)):
            c = Context.__new__(Context)
            c.lib, c.h, c.device = self.lib, C.c_void_p(self.lib.nqe_multi_ctx(h, i)), devices[i]
            self.members.append(c)

    def check(self, rc: int):
        if rc != 0:
            raise NqeError(rc, self.lib.nqe_multi_last_error(self.h).decode())

    def copy(self, table: "DeviceTable", member: int) -> "DeviceTable":
        h = C.c_void_p()
        self.check(self.lib.nqe_multi_table_copy(self.h, table.h, member, C.byref(h)))
        return DeviceTable(self.members[member], h, table.names)

    def _handles(self, tables):
        return (C.c_void_p * len(self.members))(*[t.h if t is not None else None for t in tables])

    def join_aggregate(self, left: "DeviceTable", right: Sequence[Optional["DeviceTable"]], left_key: int, right_key: int,
                       group_column: int, aggs: Sequence[tuple], names: Sequence[str]) -> "DeviceTable":
        """broadcast-build join -> group-by: `left` on any member, right[i] = member i's probe shard (or None);
        aggs = [(op, column of the joined schema)]; the result lives on member 0"""
        a = (_ffi.Agg * len(aggs))(*[_ffi.Agg(op, c) for op, c in aggs])
        h = C.c_void_p()
        self.check(self.lib.nqe_multi_join_aggregate(self.h, left.h, self._handles(right), left_key, right_key, group_column, a,
                                                     len(aggs), C.byref(h)))
        return DeviceTable(self.members[0], h, names)

    def hash_aggregate(self, shards: Sequence[Optional["DeviceTable"]], group_expr, aggs: Sequence[tuple],
                       names: Sequence[str]) -> "DeviceTable":
        """group-by over sharded input (shards[i] on member i): `group_expr` is a lowered _ffi.Expr"""
        a = (_ffi.Agg * len(aggs))(*[_ffi.Agg(op, c) for op, c in aggs])
        h = C.c_void_p()
        self.check(self.lib.nqe_multi_hash_aggregate(self.h, self._handles(shards), C.pointer(group_expr), a, len(aggs), C.byref(h)))
        return DeviceTable(self.members[0], h, names)

    def close(self):
        if self.h:
            self.lib.nqe_multi_destroy(self.h)
            self.h = None
            for c in self.members:
                c.h = None


def _normalise(arr: pa.Array) -> pa.Array:
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    if arr.offset != 0:
        arr = pa.concat_arrays([arr.slice(0, 0), arr])  # re-materialise at offset 0
        if arr.offset != 0:
            arr = pa.array(arr.to_pylist(), type=arr.type)
    return arr


class DeviceTable:
    """A RecordBatch resident in HBM (nqe_table) plus its field names."""

    def __init__(self, ctx: Context, handle: C.c_void_p, names: Sequence[str], keepalive=None):
        self.ctx = ctx
        self.h = handle
        self.names = list(names)
        self._keep = keepalive

    # ---- construction ----------------------------------------------------
    @classmethod
    def from_arrow(cls, batch, ctx: Optional[Context] = None) -> "DeviceTable":
        """Upload a pyarrow RecordBatch / Table (pin + DMA, nqe_table_upload)."""
        ctx = ctx or Context.default()
        if isinstance(batch, pa.Table):
            batch = batch.combine_chunks()
            arrays = [_normalise(batch.column(i)) for i in range(batch.num_columns)]
        else:
            arrays = [_normalise(batch.column(i)) for i in range(batch.num_columns)]
        names = list(batch.schema.names)
        n = len(arrays)
        descs = (ColumnDesc * max(n, 1))()
        for i, a in enumerate(arrays):
            dt = nqe_dtype(a.type)
            bufs = a.buffers()
            d = descs[i]
            d.dtype = dt
            d.length = len(a)
            d.null_count = a.null_count
            d.validity = bufs[0].address if (bufs[0] is not None and a.null_count) else None
            if dt == UTF8:
                d.values = bufs[1].address if bufs[1] is not None else None
                d.data = bufs[2].address if len(bufs) > 2 and bufs[2] is not None else None
                d.data_bytes = bufs[2].size if len(bufs) > 2 and bufs[2] is not None else 0
                if d.values is None:  # empty array
                    zero = np.zeros(1, dtype=np.int32)
                    arrays.append(zero)
                    d.values = zero.ctypes.data
            else:
                d.values = bufs[1].address if bufs[1] is not None else None
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_table_upload(ctx.h, descs, n, C.byref(h)))
        return cls(ctx, h, names)

    @classmethod
    def from_device_pointers(cls, ctx: Context, names: Sequence[str], dtypes: Sequence[int], ptrs: Sequence[int],
                             nrows: int, keepalive=None, validity: Optional[Sequence[int]] = None,
                             null_counts: Optional[Sequence[int]] = None) -> "DeviceTable":
        """Wrap 8-byte device columns owned by the caller (e.g. torch tensors); `validity[i]` is the device address
        of column i's LSB-first bitmap (0 = NULL-free column)."""
        n = len(ptrs)
        descs = (ColumnDesc * max(n, 1))()
        for i in range(n):
            descs[i].dtype = dtypes[i]
            descs[i].length = nrows
            descs[i].null_count = 0
            descs[i].values = ptrs[i]
            if validity is not None and validity[i]:
                descs[i].validity = validity[i]
                descs[i].null_count = null_counts[i] if null_counts is not None else -1
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_table_from_device(ctx.h, descs, n, C.byref(h)))
        return cls(ctx, h, names, keepalive)

    # ---- inspection ------------------------------------------------------
    @property
    def num_rows(self) -> int:
        return self.ctx.lib.nqe_table_num_rows(self.h)

    @property
    def num_columns(self) -> int:
        return self.ctx.lib.nqe_table_num_columns(self.h)

    def column_desc(self, i: int) -> ColumnDesc:
        d = ColumnDesc()
        self.ctx.check(self.ctx.lib.nqe_table_column(self.h, i, C.byref(d)))
        return d

    def dtypes(self) -> List[int]:
        return [self.column_desc(i).dtype for i in range(self.num_columns)]

    def schema(self) -> pa.Schema:
        return pa.schema([pa.field(nm, _NQE_TO_PA[dt]) for nm, dt in zip(self.names, self.dtypes())])

    # ---- download --------------------------------------------------------
    def to_arrow(self) -> pa.RecordBatch:
        n = self.num_rows
        arrays = []
        lib = self.ctx.lib
        for i in range(self.num_columns):
            d = self.column_desc(i)
            t = _NQE_TO_PA[d.dtype]
            has_valid = bool(d.validity) and d.null_count != 0
            vbytes = (n + 7) // 8
            valid = np.empty(max(vbytes, 1), dtype=np.uint8) if has_valid else None
            if d.dtype == UTF8:
                offs = np.empty(n + 1, dtype=np.int32)
                data = np.empty(max(d.data_bytes, 1), dtype=np.uint8)
                self.ctx.check(lib.nqe_table_download_column(
                    self.ctx.h, self.h, i, offs.ctypes.data, offs.nbytes,
                    valid.ctypes.data if has_valid else None, valid.nbytes if has_valid else 0,
                    data.ctypes.data, data.nbytes))
                bufs = [pa.py_buffer(valid) if has_valid else None, pa.py_buffer(offs), pa.py_buffer(data[:d.data_bytes])]
            else:
                nbytes = vbytes if d.dtype == BOOL else n * 8
                vals = np.empty(max(nbytes, 1), dtype=np.uint8)
                self.ctx.check(lib.nqe_table_download_column(
                    self.ctx.h, self.h, i, vals.ctypes.data, vals.nbytes,
                    valid.ctypes.data if has_valid else None, valid.nbytes if has_valid else 0, None, 0))
                bufs = [pa.py_buffer(valid) if has_valid else None, pa.py_buffer(vals[:nbytes])]
            arrays.append(pa.Array.from_buffers(t, n, bufs, null_count=d.null_count if has_valid else 0))
        return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(
            [pa.field(nm, a.type) for nm, a in zip(self.names, arrays)]))

    def slice(self, offset: int, length: int) -> "DeviceTable":
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.nqe_table_slice(self.ctx.h, self.h, offset, length, C.byref(h)))
        return DeviceTable(self.ctx, h, self.names)

    @classmethod
    def concat(cls, tables: Sequence["DeviceTable"]) -> "DeviceTable":
        """concat_batches (hash_join.rs:258-273) on the device: the rows of `tables` in order, as one new table."""
        ctx = tables[0].ctx
        arr = (C.c_void_p * len(tables))(*[t.h for t in tables])
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_table_concat(ctx.h, arr, len(tables), C.byref(h)))
        return cls(ctx, h, list(tables[0].names))

    def free(self):
        if self.h:
            if self.ctx.h:  # a table that outlives its (closed) context is dropped, not freed through a dangling nqe_ctx
                self.ctx.lib.nqe_table_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
