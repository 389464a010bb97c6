"""world_size-2 gloo tests of the multi-GPU plans (distributed.py) on CPU: shuffled and broadcast join + group-by,
the stand-alone distributed group-by and the stand-alone distributed hash join (both plans).

The exchange orchestration (radix partition -> all-to-all with uneven splits / all-gather of the build side ->
local operator -> partial states exchanged by group-key radix -> merge -> gather) runs for real over gloo; the
per-rank operator calls are answered by the CPU oracle (test infrastructure) instead of
the CUDA kernels, and the result must equal the single-process oracle's."""
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


    def partial_aggregate(self, cols):
        O = self.O
        b = O.Batch(["k", "v"], [O.Col("i64", cols[0].numpy()), O.Col("f64", cols[1].numpy().view(np.float64))])
        a = O.aggregate(b, ("col", 0), [("count", 1), ("sum", 1), ("min", 1), ("max", 1), ("min", 0)])
        return [torch.from_numpy(a.cols[4].values.astype(np.int64)), torch.from_numpy(a.cols[0].values.astype(np.int64)),
                torch.from_numpy(a.cols[1].values.view(np.int64).copy()),
                torch.from_numpy(a.cols[2].values.view(np.int64).copy()),
                torch.from_numpy(a.cols[3].values.view(np.int64).copy())]

    def hash_join(self, lcols, rcols):
        O = self.O
        L = O.Batch([f"l{i}" for i in range(len(lcols))], [O.Col("i64", c.numpy()) for c in lcols])
        R = O.Batch([f"r{i}" for i in range(len(rcols))], [O.Col("i64", c.numpy()) for c in rcols])
        j = O.hash_join_c(L, R, 0, 0)
        return [torch.from_numpy(np.ascontiguousarray(c.values.astype(np.int64))) for c in j.cols]


N_BUILD, N_PROBE, GROUPS = 4000, 30000, 97


def _tables(start_l, n_l, start_r, n_r):
    from oracle import oracle as O
    lk = O.gen_perm_i64(start_l, n_l, 7368787 % N_BUILD or 1, N_BUILD)
    la = lk % GROUPS
    fk = O.gen_mod_i64(47, start_r, n_r, int(N_BUILD * 1.25))  # some probe rows find no match
    rb = O.gen_unif_f64(48, start_r, n_r, 100.0)
    return lk, la, fk, rb


def _worker(rank, world, port, out_path, peer_dir=None, plan="shuffle"):
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
    if plan == "broadcast":
        ph = D.Phases(torch)
        merged, sent = D.broadcast_join_group_by(dist, torch, engine, lcols, rcols, world, phases=ph)
        assert {"build_all_gather", "local_join_aggregate", "partial_exchange", "merge"} <= set(ph.ms())
    else:
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


@pytest.mark.parametrize("plan", ["shuffle", "shuffle-peer", "broadcast"])
def test_join_group_by_world2(tmp_path, plan):
    from oracle import oracle as O
    world = 2
    peer = plan == "shuffle-peer"
    out = str(tmp_path / "merged.npz")
    port = 29500 + (os.getpid() % 2000) + ["shuffle", "shuffle-peer", "broadcast"].index(plan)
    mp.spawn(_worker, args=(world, port, out, str(tmp_path) if peer else None, plan.split("-")[0]), nprocs=world, join=True)
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


def _worker_ops(rank, world, port, out_dir):
    """stand-alone distributed group-by and hash join (both plans)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    D = import_module("naive-query-engine_b200.distributed")
    nl, nr = N_BUILD // world, N_PROBE // world
    lk, la, fk, rb = _tables(rank * nl, nl, rank * nr, nr)
    engine = OracleEngine()
    # group by fk % 301 over the probe table
    gk = torch.from_numpy((fk % 301).astype(np.int64))
    gv = torch.from_numpy(rb.view(np.int64).copy())
    full, sent_g = D.distributed_group_by(dist, torch, engine, [gk, gv], world, gather=True)
    own, _ = D.distributed_group_by(dist, torch, engine, [gk, gv], world, gather=False)
    owner = (_mix64(own[0].numpy().view(np.uint64)) % np.uint64(world)).astype(np.int64)
    assert np.all(owner == rank)  # without the gather every rank keeps exactly the groups it owns
    lcols = [torch.from_numpy(lk), torch.from_numpy(la)]
    rcols = [torch.from_numpy(fk), torch.from_numpy(rb.view(np.int64).copy())]
    jb, wire_b = D.distributed_hash_join(dist, torch, engine, lcols, rcols, world, plan="broadcast")
    js, wire_s = D.distributed_hash_join(dist, torch, engine, lcols, rcols, world, plan="shuffle")
    np.savez(os.path.join(out_dir, f"ops_{rank}.npz"), gb=np.stack([c.numpy() for c in full]), sent_g=sent_g,
             jb=np.stack([c.numpy() for c in jb]), js=np.stack([c.numpy() for c in js]), wire_b=wire_b, wire_s=wire_s)
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_group_by_and_hash_join_world2(tmp_path):
    from oracle import oracle as O
    world = 2
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_worker_ops, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(str(tmp_path / f"ops_{r}.npz")) for r in range(world)]
    lk, la, fk, rb = _tables(0, N_BUILD, 0, N_PROBE)
    # ---- group-by: every rank holds the whole result
    T = O.Batch(["k", "v"], [O.Col("i64", (fk % 301).astype(np.int64)), O.Col("f64", rb)])
    want = O.aggregate(T, ("col", 0), [("min", 0), ("count", 1), ("sum", 1), ("min", 1), ("max", 1)])
    wk = want.cols[0].values.astype(np.int64)
    wo = np.argsort(wk)
    for r in range(world):
        gb = got[r]["gb"]
        go = np.argsort(gb[0])
        assert np.array_equal(wk[wo], gb[0][go]) and np.array_equal(want.cols[1].values[wo].astype(np.int64), gb[1][go])
        assert np.allclose(want.cols[2].values[wo], gb[2][go].view(np.float64), rtol=1e-9, atol=0)
        assert np.array_equal(want.cols[3].values[wo], gb[3][go].view(np.float64))
        assert np.array_equal(want.cols[4].values[wo], gb[4][go].view(np.float64))
    assert sum(int(g["sent_g"]) for g in got) > 0
    # ---- hash join
    L = O.Batch(["k", "a"], [O.Col("i64", lk), O.Col("i64", la)])
    R = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("i64", rb.view(np.int64).copy())])
    wj = np.stack([c.values.astype(np.int64) for c in O.hash_join_c(L, R, 0, 0).cols])
    # broadcast plan: concatenation in rank order IS the reference's probe-row-major order
    assert np.array_equal(np.concatenate([g["jb"] for g in got], axis=1), wj)
    # shuffle plan: same multiset of rows
    js = np.concatenate([g["js"] for g in got], axis=1)
    assert js.shape == wj.shape
    assert np.array_equal(js[:, np.lexsort(js[::-1])], wj[:, np.lexsort(wj[::-1])])
    assert all(int(g["wire_b"]) > 0 and int(g["wire_s"]) > 0 for g in got)
