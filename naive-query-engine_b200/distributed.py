"""Multi-GPU plans for hash-join and group-by: one process per GPU,
`torch.distributed` for the plumbing.

Three operators, each with the exchange SURVEY.md 8(e) names for it:

    distributed_group_by     local pre-aggregate -> partial states exchanged by group-key radix -> merge
                             (plan (ii) of 8e: G x 40 bytes on the wire instead of N x 16)
    distributed_hash_join    `shuffle`: both sides exchanged by join-key radix, local join of what arrives
                             `broadcast`: the build side all-gathered, every rank joins its own probe rows
                             (output = concatenation in rank order = the reference's probe-row-major order)
    join + group-by          `shuffled_join_group_by` (both tables shuffled by join key, below) or
                             `broadcast_join_group_by` (build side all-gathered -- 160 MB for configs[4] -- the probe
                             rows never move, only partial states are exchanged; the default for small build sides)

The shuffle in detail:

    rank r owns a row range of L (build) and R (probe)
    1+2. radix partition fused with the exchange (`exchange_peer`): every rank counts its rows per
       destination (nqe_partition_counts), the ranks all-gather those P x P counts, and one scatter
       kernel (nqe_shuffle_scatter) stores each row straight into the destination GPU's receive
       buffer through peer-mapped memory (NVLink 5 / NVSwitch; symmetric memory for the mapping)
       -- no partitioned staging copy, no separate all-to-all.
       Baseline kept for comparison and for engines without peer memory (`exchange`):
       nqe_radix_partition then dist.all_to_all_single with uneven splits (NCCL).
    3. local fused join -> group-by on the received rows   (nqe_join_aggregate; min(group) carries the key,
                                                            because the reference's aggregate emits no key column)
    4. all-gather the partial states and merge them        (count -> sum of counts, sum -> sum, min, max;
                                                            avg = sum / count)

Filter / projection need no exchange: every rank runs them on its own rows.

The orchestration is written against a small `engine` interface so that the CPU
tests can drive it over gloo with a stand-in engine, while bench.py drives it over
NCCL with the CUDA engine below.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple


class Engine:
    """What the exchange plan needs from the execution layer.  Columns are 1-D
    torch tensors (int64 storage; float64 columns are bit-cast) on the engine's device."""

    def partition(self, cols: Sequence, key: int, parts: int) -> Tuple[List, List[int]]:
        """-> (columns permuted so that rows of destination p are contiguous, counts[parts])"""
        raise NotImplementedError

    def partition_counts(self, cols: Sequence, key: int, parts: int) -> List[int]:
        """rows of `cols` per destination (dest = mix64(key) % parts)"""
        raise NotImplementedError

    def alloc_exchange(self, capacity_rows: int, n_cols: int, group):
        """receive buffers every rank of `group` can store into: .capacity, .local(c) -> this rank's column
        c buffer (tensor), .barrier() -> all ranks' earlier work on the buffers is complete"""
        raise NotImplementedError

    def scatter_to_peers(self, cols: Sequence, key: int, parts: int, xbuf, offsets: Sequence[int]) -> None:
        """store every row into column buffers of rank mix64(key) % parts, starting at row offsets[dest]"""
        raise NotImplementedError

    def join_partial_aggregate(self, lcols: Sequence, rcols: Sequence):
        """inner join L(k, a) x R(fk, b) on k = fk, group by a ->
        columns [key(a) i64, count i64, sum f64-bits, min f64-bits, max f64-bits]"""
        raise NotImplementedError

    def merge_partials(self, cols: Sequence):
        """group by key over partial states -> [key, count, sum, min, max] (one row per key)"""
        raise NotImplementedError

    def partial_aggregate(self, cols: Sequence):
        """group by cols[0] over (k i64, v f64-bits) -> [key, count, sum f64-bits, min f64-bits, max f64-bits]"""
        raise NotImplementedError

    def hash_join(self, lcols: Sequence, rcols: Sequence):
        """inner join on lcols[0] = rcols[0] (Int64 keys) -> columns of L then columns of R, probe-row-major"""
        raise NotImplementedError


def exchange(dist, torch, engine: Engine, cols: Sequence, key: int, world: int):
    """Steps 1-2 for one table.  Returns (received columns, rows sent to other ranks)."""
    part, counts = engine.partition(cols, key, world)
    rank = dist.get_rank()
    send = torch.tensor(counts, dtype=torch.int64, device=part[0].device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    send_l = [int(x) for x in counts]
    recv_l = [int(x) for x in recv.tolist()]
    total = sum(recv_l)
    outs = []
    for c in part:
        dst = torch.empty(total, dtype=c.dtype, device=c.device)
        dist.all_to_all_single(dst, c, output_split_sizes=recv_l, input_split_sizes=send_l)
        outs.append(dst)
    return outs, sum(send_l) - send_l[rank]


def peer_offsets(count_matrix: Sequence[Sequence[int]], rank: int):
    """count_matrix[src][dst] = rows src sends to dst.  -> (first row `rank` writes in every destination's
    receive buffer, rows `rank` receives in total, the largest receive total over all ranks)."""
    world = len(count_matrix)
    offsets = [sum(count_matrix[s][d] for s in range(rank)) for d in range(world)]
    totals = [sum(count_matrix[s][d] for s in range(world)) for d in range(world)]
    return offsets, totals[rank], max(totals)


def exchange_peer(dist, torch, engine: Engine, cols: Sequence, key: int, world: int, xbuf):
    """Steps 1-2 fused: rows go straight into the destination ranks' receive buffers (`xbuf`, from
    engine.alloc_exchange).  Returns (received columns -- views of this rank's buffer, rows sent to others)."""
    rank = dist.get_rank()
    counts = engine.partition_counts(cols, key, world)
    mine = torch.tensor(counts, dtype=torch.int64, device=cols[0].device)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    matrix = [[int(x) for x in t.tolist()] for t in allc]
    offsets, total, need = peer_offsets(matrix, rank)
    if need > xbuf.capacity:
        raise RuntimeError(f"exchange buffer too small: {need} rows needed, capacity {xbuf.capacity}")
    xbuf.barrier()  # every rank has finished reading what the previous exchange delivered
    engine.scatter_to_peers(cols, key, world, xbuf, offsets)
    xbuf.barrier()  # every rank's stores have landed
    return [xbuf.local(c)[:total] for c in range(len(cols))], sum(counts) - counts[rank]


class Phases:
    """Per-phase device times of one plan execution: `mark(name)` closes the phase that started at the previous mark.
    Events are recorded on the current stream of `torch`; `ms()` synchronises and returns {phase: milliseconds}."""

    def __init__(self, torch):
        self.torch, self.names, self.events = torch, [], []
        self._rec()

    def _rec(self):
        cuda = getattr(self.torch, "cuda", None)
        if cuda is not None and cuda.is_available():
            e = cuda.Event(enable_timing=True)
            e.record()
            self.events.append(e)
        else:
            import time
            self.events.append(time.perf_counter())

    def mark(self, name: str):
        self.names.append(name)
        self._rec()

    def ms(self):
        out = {}
        for i, n in enumerate(self.names):
            a, b = self.events[i], self.events[i + 1]
            if isinstance(a, float):
                out[n] = out.get(n, 0.0) + (b - a) * 1e3
            else:
                b.synchronize()
                out[n] = out.get(n, 0.0) + a.elapsed_time(b)
        return out


def all_gather_sizes(dist, torch, n: int, world: int, device):
    """every rank's row count, in rank order (one small collective + one host synchronisation)"""
    sizes = torch.tensor([n], dtype=torch.int64, device=device)
    all_sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes)
    return [int(x) for x in all_sizes.tolist()]


def all_gather_rows(dist, torch, cols: Sequence, world: int, nl: Optional[Sequence[int]] = None):
    """All ranks' rows of `cols` (equal-dtype 1-D tensors) in rank order, with ONE collective per call for the sizes
    (skipped when the caller already has them: `nl`) and one for the data: the columns travel packed as a
    (rows, n_cols) matrix, padded to the largest rank."""
    n = int(cols[0].numel())
    if nl is None:
        nl = all_gather_sizes(dist, torch, n, world, cols[0].device)
    nmax = max(max(nl), 1)
    packed = torch.zeros((nmax, len(cols)), dtype=cols[0].dtype, device=cols[0].device)
    packed[:n] = torch.stack(list(cols), dim=1)
    out = torch.empty((world * nmax, len(cols)), dtype=cols[0].dtype, device=cols[0].device)
    dist.all_gather_into_tensor(out, packed)
    if all(x == nmax for x in nl):
        rows = out
    else:
        rows = torch.cat([out[r * nmax:r * nmax + nl[r]] for r in range(world)])
    return [rows[:, c].contiguous() for c in range(len(cols))], nl


def exchange_partials(dist, torch, engine: Engine, partial: Sequence, world: int):
    """Partial aggregate states [key, count, sum, min, max] -> every rank receives the states of the groups it owns
    (owner = mix64(key) % world): one radix partition of the (small) partial table, one count exchange, ONE
    all-to-all of 40-byte records.  Returns (received columns, rows sent to other ranks)."""
    part, counts = engine.partition(partial, 0, world)
    rank = dist.get_rank()
    dev = part[0].device
    send = torch.tensor(counts, dtype=torch.int64, device=dev)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    send_l, recv_l = [int(x) for x in counts], [int(x) for x in recv.tolist()]
    rows = torch.stack([c if c.dtype == torch.int64 else c.view(torch.int64) for c in part], dim=1).contiguous()
    got = torch.empty((sum(recv_l), len(part)), dtype=torch.int64, device=dev)
    dist.all_to_all_single(got, rows, output_split_sizes=recv_l, input_split_sizes=send_l)
    return [got[:, c].contiguous() for c in range(len(part))], sum(send_l) - send_l[rank]


GATHER_MERGE_MAX_ROWS = 1 << 22  # partial rows over all ranks up to which "all-gather, merge everywhere" beats the radix exchange


def merge_exchanged_partials(dist, torch, engine: Engine, partial: Sequence, world: int, gather: bool = True, phases=None):
    """Step 4 of every aggregate plan: partial states -> final groups.  The states are exchanged by group-key radix
    and merged by their owners; gather=True then all-gathers the owners' slices so that every rank returns the whole
    result [key, count, sum, min, max], else every rank returns the groups it owns.
    When the whole result is wanted everywhere and the partial tables are small (G x 40 bytes x world: 32 MB for
    configs[4] on 8 GPUs) the radix step is skipped: one all-gather of the partial states, every rank merges all of
    them -- two collectives and one host synchronisation instead of five and three (measured at 8 GPUs: 1.28 -> ms
    in bench.py's phase table)."""
    nl = all_gather_sizes(dist, torch, int(partial[0].numel()), world, partial[0].device) if gather else None
    if gather and sum(nl) <= GATHER_MERGE_MAX_ROWS:  # decided from the gathered sizes: every rank takes the same branch
        rows, nl = all_gather_rows(dist, torch, [c if c.dtype == torch.int64 else c.view(torch.int64) for c in partial], world, nl)
        if phases is not None:
            phases.mark("partial_exchange")
        merged = engine.merge_partials(rows)
        if phases is not None:
            phases.mark("merge")
        return merged, sum(nl) - nl[dist.get_rank()]
    recv, sent = exchange_partials(dist, torch, engine, partial, world)
    if phases is not None:
        phases.mark("partial_exchange")
    merged = engine.merge_partials(recv)
    if phases is not None:
        phases.mark("merge")
    if gather:
        merged, _ = all_gather_rows(dist, torch, merged, world)
        if phases is not None:
            phases.mark("result_gather")
    return merged, sent


def shuffled_join_group_by(dist, torch, engine: Engine, lcols: Sequence, rcols: Sequence, world: int, xbufs=None,
                           phases=None):
    """Steps 1-4.  Every rank returns the full merged result
    [key, count, sum, min, max] (tensors) and the number of rows it sent over the wire.
    xbufs = (build-side, probe-side) exchange buffers: rows travel through peer memory; None: NCCL all-to-all."""
    if xbufs is not None:
        l_recv, s1 = exchange_peer(dist, torch, engine, lcols, 0, world, xbufs[0])
        r_recv, s2 = exchange_peer(dist, torch, engine, rcols, 0, world, xbufs[1])
    else:
        l_recv, s1 = exchange(dist, torch, engine, lcols, 0, world)
        r_recv, s2 = exchange(dist, torch, engine, rcols, 0, world)
    if phases is not None:
        phases.mark("row_exchange")
    partial = engine.join_partial_aggregate(l_recv, r_recv)
    if phases is not None:
        phases.mark("local_join_aggregate")
    merged, _ = merge_exchanged_partials(dist, torch, engine, partial, world, True, phases)
    return merged, s1 + s2


def broadcast_join_group_by(dist, torch, engine: Engine, lcols: Sequence, rcols: Sequence, world: int, phases=None):
    """Join + group-by WITHOUT moving the probe side: the build side (small: 160 MB for configs[4]) is all-gathered,
    every rank runs the fused join -> partial aggregate over its own probe rows against the whole build table, and only
    partial states (G x 40 bytes) are exchanged and merged.  Every rank returns ([key, count, sum, min, max], rows
    received over the wire)."""
    rank = dist.get_rank()
    l_all, nl = all_gather_rows(dist, torch, lcols, world)
    if phases is not None:
        phases.mark("build_all_gather")
    partial = engine.join_partial_aggregate(l_all, rcols)
    if phases is not None:
        phases.mark("local_join_aggregate")
    merged, _ = merge_exchanged_partials(dist, torch, engine, partial, world, True, phases)
    return merged, sum(nl) - nl[rank]


def distributed_group_by(dist, torch, engine: Engine, cols: Sequence, world: int, gather: bool = True, phases=None):
    """`select count(v), sum(v), avg(v), min(v), max(v) from t group by k` over row-sharded t(k, v): local
    pre-aggregate -> partial states exchanged by group-key radix -> merge (plan (ii) of SURVEY.md 8e; avg = sum / count).
    -> ([key, count, sum, min, max], partial rows sent to other ranks)."""
    partial = engine.partial_aggregate(cols)
    if phases is not None:
        phases.mark("local_aggregate")
    return merge_exchanged_partials(dist, torch, engine, partial, world, gather, phases)


def distributed_hash_join(dist, torch, engine: Engine, lcols: Sequence, rcols: Sequence, world: int, plan: str = "broadcast",
                          xbufs=None, phases=None):
    """`select * from L join R on L.c0 = R.c0` over row-sharded tables; every rank keeps its slice of the joined rows
    (columns of L then columns of R, nothing is gathered).
      broadcast: L all-gathered, local join with this rank's R rows -- the concatenation of the ranks' outputs in rank
                 order is exactly the reference's probe-row-major order;
      shuffle:   both sides exchanged by join-key radix (through `xbufs` peer memory if given, else NCCL all-to-all),
                 local join of what arrives -- same multiset of rows, order by (key owner, arrival).
    -> (joined columns, rows received or sent over the wire)."""
    rank = dist.get_rank()
    if plan == "broadcast":
        l_all, nl = all_gather_rows(dist, torch, lcols, world)
        if phases is not None:
            phases.mark("build_all_gather")
        out = engine.hash_join(l_all, rcols)
        wire = sum(nl) - nl[rank]
    elif plan == "shuffle":
        if xbufs is not None:
            l_recv, s1 = exchange_peer(dist, torch, engine, lcols, 0, world, xbufs[0])
            r_recv, s2 = exchange_peer(dist, torch, engine, rcols, 0, world, xbufs[1])
        else:
            l_recv, s1 = exchange(dist, torch, engine, lcols, 0, world)
            r_recv, s2 = exchange(dist, torch, engine, rcols, 0, world)
        if phases is not None:
            phases.mark("row_exchange")
        out = engine.hash_join(l_recv, r_recv)
        wire = s1 + s2
    else:
        raise ValueError(plan)
    if phases is not None:
        phases.mark("local_join")
    return out, wire


class CudaEngine(Engine):
    """The CUDA execution layer behind the Engine interface (through the C ABI)."""

    I64, F64 = 2, 4

    def __init__(self, nq, ctx, torch):
        self.nq, self.ctx, self.torch = nq, ctx, torch

    def _bind(self):
        """Stream contract: every kernel of this engine runs on torch's CURRENT stream (the context is re-bound to it
        at the start of each call), which is also the stream the symmetric-memory barriers and NCCL collectives of
        the plans above are queued on -- so a peer's stores, the barrier after them and the kernels that read the
        receive buffer are ordered by the stream itself, with no host synchronisation in between."""
        self.ctx.set_stream(self.torch.cuda.current_stream().cuda_stream)

    def _table(self, names, dtypes, cols):
        return self.nq.DeviceTable.from_device_pointers(self.ctx, names, dtypes, [c.data_ptr() for c in cols],
                                                        int(cols[0].numel()), keepalive=list(cols))

    def _column(self, table, i, n, copy=True):
        if n == 0:
            return self.torch.empty(0, dtype=self.torch.int64, device="cuda")
        d = table.column_desc(i)

        class _View:  # __cuda_array_interface__ over the table's device buffer
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (d.values, False), "version": 2}
        v = self.torch.as_tensor(_View(), device="cuda")
        return v.clone() if copy else v

    def partition(self, cols, key, parts):
        self._bind()
        ctx = self.ctx
        t = self._table([f"c{i}" for i in range(len(cols))], [self.I64] * len(cols), cols)
        h = C.c_void_p()
        counts = (C.c_int64 * parts)()
        ctx.check(ctx.lib.nqe_radix_partition(ctx.h, t.h, key, parts, C.byref(h), counts))
        out = self.nq.DeviceTable(ctx, h, t.names)
        n = int(cols[0].numel())
        res = [self._column(out, i, n, copy=False) for i in range(len(cols))]
        for r in res:
            r._nqe_keepalive = out  # the views borrow the partitioned table's buffers
        t.free()
        return res, [int(x) for x in counts]

    def partition_counts(self, cols, key, parts):
        self._bind()
        ctx = self.ctx
        t = self._table([f"c{i}" for i in range(len(cols))], [self.I64] * len(cols), cols)
        counts = (C.c_int64 * parts)()
        ctx.check(ctx.lib.nqe_partition_counts(ctx.h, t.h, key, parts, counts))
        t.free()
        return [int(x) for x in counts]

    class _SymmExchange:
        """Per-column symmetric-memory receive buffers (torch.distributed._symmetric_memory does the
        CUDA IPC mapping and the stream barrier: plumbing); the stores are ours (nqe_shuffle_scatter)."""

        def __init__(self, torch, capacity_rows, n_cols, group):
            import torch.distributed._symmetric_memory as symm_mem
            self.capacity = int(capacity_rows)
            self.tensors = [symm_mem.empty(self.capacity, dtype=torch.int64, device="cuda") for _ in range(n_cols)]
            self.handles = [symm_mem.rendezvous(t, group) for t in self.tensors]
            self.ptrs = [[int(p) for p in h.buffer_ptrs] for h in self.handles]  # [col][rank]

        def local(self, c):
            return self.tensors[c]

        def barrier(self):
            self.handles[0].barrier(channel=0)

    def alloc_exchange(self, capacity_rows, n_cols, group):
        return self._SymmExchange(self.torch, capacity_rows, n_cols, group)

    def scatter_to_peers(self, cols, key, parts, xbuf, offsets):
        self._bind()
        ctx = self.ctx
        t = self._table([f"c{i}" for i in range(len(cols))], [self.I64] * len(cols), cols)
        dst = (C.c_void_p * (len(cols) * parts))(*[xbuf.ptrs[c][p] for c in range(len(cols)) for p in range(parts)])
        offs = (C.c_int64 * parts)(*[int(o) for o in offsets])
        ctx.check(ctx.lib.nqe_shuffle_scatter(ctx.h, t.h, key, parts, dst, offs))
        t.free()

    def join_partial_aggregate(self, lcols, rcols):
        self._bind()
        ctx, nq = self.ctx, self.nq
        L = self._table(["k", "a"], [self.I64, self.I64], lcols)
        R = self._table(["fk", "b"], [self.I64, self.F64], rcols)
        # join output schema: k, a, fk, b -> count(b), sum(b), min(b), max(b), key group by a  (op 5 = NQE_AGG_GROUP_KEY:
        # the reference's aggregate emits no key column; the merge needs it)
        aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(op, c) for op, c in [(0, 3), (1, 3), (3, 3), (4, 3), (5, 0)]])
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_join_aggregate(ctx.h, L.h, R.h, 0, 0, 1, aggs, 5, C.byref(h)))
        part = nq.DeviceTable(ctx, h, ["count", "sum", "min", "max", "key"])
        g = part.num_rows
        cols = [self._column(part, i, g) for i in range(5)]
        self.torch.cuda.current_stream().synchronize()
        part.free(); L.free(); R.free()
        return [cols[4], cols[0], cols[1], cols[2], cols[3]]

    def merge_partials(self, cols):
        self._bind()
        ctx, nq, torch = self.ctx, self.nq, self.torch
        key, cnt, s, mn, mx = cols
        cnt_f = cnt.to(torch.float64).view(torch.int64)  # counts are summed as f64 (exact below 2^53)
        M = self._table(["key", "cnt", "sum", "min", "max"], [self.I64, self.F64, self.F64, self.F64, self.F64],
                        [key, cnt_f, s, mn, mx])
        aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(op, c) for op, c in [(5, 0), (1, 1), (1, 2), (3, 3), (4, 4)]])
        ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(M.names)
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, M.h, C.pointer(ke), aggs, 5, C.byref(h)))
        out = nq.DeviceTable(ctx, h, ["key", "count", "sum", "min", "max"])
        g = out.num_rows
        res = [self._column(out, i, g) for i in range(5)]
        torch.cuda.current_stream().synchronize()
        out.free(); M.free()
        res[1] = res[1].view(torch.float64).to(torch.int64)
        return res

    def partial_aggregate(self, cols):
        self._bind()
        ctx, nq, torch = self.ctx, self.nq, self.torch
        T = self._table(["k", "v"], [self.I64, self.F64], cols)
        aggs = (nq._ffi.Agg * 5)(*[nq._ffi.Agg(op, c) for op, c in [(0, 1), (1, 1), (3, 1), (4, 1), (5, 0)]])
        ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(T.names)
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_aggregate(ctx.h, T.h, C.pointer(ke), aggs, 5, C.byref(h)))
        part = nq.DeviceTable(ctx, h, ["count", "sum", "min", "max", "key"])
        g = part.num_rows
        res = [self._column(part, i, g) for i in range(5)]
        torch.cuda.current_stream().synchronize()
        part.free(); T.free()
        return [res[4], res[0], res[1], res[2], res[3]]

    def hash_join(self, lcols, rcols):
        self._bind()
        ctx, nq, torch = self.ctx, self.nq, self.torch
        L = self._table([f"l{i}" for i in range(len(lcols))], [self.I64] * len(lcols), lcols)
        R = self._table([f"r{i}" for i in range(len(rcols))], [self.I64] * len(rcols), rcols)
        h = C.c_void_p()
        ctx.check(ctx.lib.nqe_hash_join(ctx.h, L.h, R.h, 0, 0, C.byref(h)))
        out = nq.DeviceTable(ctx, h, L.names + R.names)
        n = out.num_rows
        res = [self._column(out, i, n) for i in range(len(lcols) + len(rcols))]
        torch.cuda.current_stream().synchronize()
        out.free(); L.free(); R.free()
        return res
