"""world_size-2 gloo test of the multi-GPU join + group-by plan (distributed.py) on CPU.

The exchange orchestration (radix partition -> all-to-all with uneven splits ->
local join + partial aggregate -> all-gather + merge) runs for real over gloo; the
per-rank operator calls are answered by the CPU oracle (test infrastructure) instead of
the CUDA kernels, and the merged result must equal the single-process oracle's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _mix64(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


class OracleEngine:
    def __init__(self):
        from importlib import import_module
        from oracle import oracle as O
        self.O = O
        self.Engine = import_module("naive-query-engine_b200.distributed").Engine

    def partition(self, cols, key, parts):
        k = cols[key].numpy()
        dest = (_mix64(k.view(np.uint64)) % np.uint64(parts)).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=parts).tolist()
        return [torch.from_numpy(c.numpy()[order].copy()) for c in cols], counts

    # ---- peer-memory exchange, emulated with files every rank maps (np.memmap) ----
    def partition_counts(self, cols, key, parts):
        k = cols[key].numpy()
        dest = (_mix64(k.view(np.uint64)) % np.uint64(parts)).astype(np.int64)
        return np.bincount(dest, minlength=parts).tolist()

    def alloc_exchange(self, capacity_rows, n_cols, group, shared_dir=None, tag="x"):
        rank, world = dist.get_rank(), dist.get_world_size()

        class X:
            pass
        x = X()
        x.capacity = capacity_rows
        for c in range(n_cols):  # every rank creates its own receive files, then maps everybody's
            np.memmap(os.path.join(shared_dir, f"{tag}_{rank}_{c}.bin"), dtype=np.int64, mode="w+", shape=(capacity_rows,)).flush()
        dist.barrier()
        x.maps = [[np.memmap(os.path.join(shared_dir, f"{tag}_{p}_{c}.bin"), dtype=np.int64, mode="r+", shape=(capacity_rows,))
                   for p in range(world)] for c in range(n_cols)]
        x.local = lambda c: torch.from_numpy(np.array(x.maps[c][rank]))

        def barrier():
            for row in x.maps:
                for m in row:
                    m.flush()
            dist.barrier()
        x.barrier = barrier
        return x

    def scatter_to_peers(self, cols, key, parts, xbuf, offsets):
        k = cols[key].numpy()
        dest = (_mix64(k.view(np.uint64)) % np.uint64(parts)).astype(np.int64)
        for p in range(parts):
            sel = np.nonzero(dest == p)[0]
            for c, col in enumerate(cols):
                xbuf.maps[c][p][offsets[p]:offsets[p] + len(sel)] = col.numpy()[sel]

    def join_partial_aggregate(self, lcols, rcols):
        O = self.O
        L = O.Batch(["k", "a"], [O.Col("i64", lcols[0].numpy()), O.Col("i64", lcols[1].numpy())])
        R = O.Batch(["fk", "b"], [O.Col("i64", rcols[0].numpy()), O.Col("f64", rcols[1].numpy().view(np.float64))])
        j = O.hash_join_c(L, R, 0, 0)
        a = O.aggregate(j, ("col", 1), [("count", 3), ("sum", 3), ("min", 3), ("max", 3), ("min", 1)])
        key = a.cols[4].values.astype(np.int64)
        return [torch.from_numpy(key), torch.from_numpy(a.cols[0].values.astype(np.int64)),
                torch.from_numpy(a.cols[1].values.view(np.int64).copy()),
                torch.from_numpy(a.cols[2].values.view(np.int64).copy()),
                torch.from_numpy(a.cols[3].values.view(np.int64).copy())]

    def merge_partials(self, cols):
        O = self.O
        key, cnt, s, mn, mx = [c.numpy() for c in cols]
        b = O.Batch(["key", "cnt", "sum", "min", "max"],
                    [O.Col("i64", key), O.Col("f64", cnt.astype(np.float64)), O.Col("f64", s.view(np.float64)),
                     O.Col("f64", mn.view(np.float64)), O.Col("f64", mx.view(np.float64))])
        m = O.aggregate(b, ("col", 0), [("min", 0), ("sum", 1), ("sum", 2), ("min", 3), ("max", 4)])
        return [torch.from_numpy(m.cols[0].values.astype(np.int64)), torch.from_numpy(m.cols[1].values.astype(np.int64)),
                torch.from_numpy(m.cols[2].values.view(np.int64).copy()),
                torch.from_numpy(m.cols[3].values.view(np.int64).copy()),
                torch.from_numpy(m.cols[4].values.view(np.int64).copy())]


N_BUILD, N_PROBE, GROUPS = 4000, 30000, 97


def _tables(start_l, n_l, start_r, n_r):
    from oracle import oracle as O
    lk = O.gen_perm_i64(start_l, n_l, 7368787 % N_BUILD or 1, N_BUILD)
    la = lk % GROUPS
    fk = O.gen_mod_i64(47, start_r, n_r, int(N_BUILD * 1.25))  # some probe rows find no match
    rb = O.gen_unif_f64(48, start_r, n_r, 100.0)
    return lk, la, fk, rb


def _worker(rank, world, port, out_path, peer_dir=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    D = import_module("naive-query-engine_b200.distributed")
    nl, nr = N_BUILD // world, N_PROBE // world
    lk, la, fk, rb = _tables(rank * nl, nl, rank * nr, nr)
    lcols = [torch.from_numpy(lk), torch.from_numpy(la)]
    rcols = [torch.from_numpy(fk), torch.from_numpy(rb.view(np.int64).copy())]
    engine = OracleEngine()
    xbufs = None
    if peer_dir is not None:  # rows travel through "peer memory" (files mapped by every rank)
        xbufs = (engine.alloc_exchange(N_BUILD, 2, None, peer_dir, "l"), engine.alloc_exchange(N_PROBE, 2, None, peer_dir, "r"))
    merged, sent = D.shuffled_join_group_by(dist, torch, engine, lcols, rcols, world, xbufs)
    if rank == 0:
        np.savez(out_path, key=merged[0].numpy(), count=merged[1].numpy(), sum=merged[2].numpy().view(np.float64),
                 min=merged[3].numpy().view(np.float64), max=merged[4].numpy().view(np.float64), sent=sent)
    dist.barrier()
    dist.destroy_process_group()


def test_peer_offsets():
    from importlib import import_module
    D = import_module("naive-query-engine_b200.distributed")
    m = [[5, 1, 2], [0, 7, 3], [4, 4, 4]]  # m[src][dst]
    assert D.peer_offsets(m, 0) == ([0, 0, 0], 9, 12)
    assert D.peer_offsets(m, 1) == ([5, 1, 2], 12, 12)
    assert D.peer_offsets(m, 2) == ([5, 8, 5], 9, 12)


@pytest.mark.parametrize("peer", [False, True])
def test_shuffled_join_group_by_world2(tmp_path, peer):
    from oracle import oracle as O
    world = 2
    out = str(tmp_path / "merged.npz")
    port = 29500 + (os.getpid() % 2000) + (1 if peer else 0)
    mp.spawn(_worker, args=(world, port, out, str(tmp_path) if peer else None), nprocs=world, join=True)
    got = np.load(out)
    lk, la, fk, rb = _tables(0, N_BUILD, 0, N_PROBE)
    L = O.Batch(["k", "a"], [O.Col("i64", lk), O.Col("i64", la)])
    R = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("f64", rb)])
    want = O.aggregate(O.hash_join_c(L, R, 0, 0), ("col", 1), [("min", 1), ("count", 3), ("sum", 3), ("min", 3), ("max", 3)])
    wk = want.cols[0].values.astype(np.int64)
    wo, go = np.argsort(wk), np.argsort(got["key"])
    assert np.array_equal(wk[wo], got["key"][go])
    assert np.array_equal(want.cols[1].values[wo].astype(np.int64), got["count"][go])
    assert np.allclose(want.cols[2].values[wo], got["sum"][go], rtol=1e-9, atol=0)
    assert np.array_equal(want.cols[3].values[wo], got["min"][go])
    assert np.array_equal(want.cols[4].values[wo], got["max"][go])
    assert int(got["sent"]) > 0  # rows really crossed ranks
