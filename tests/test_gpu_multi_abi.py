"""Several GPUs behind the C ABI from ONE process (include/nqe.h nqe_multi_*, csrc/multi.cu): the broadcast-build
join -> group-by and the sharded group-by against numpy over the concatenated shards.  Members are the box's GPUs; on a
one-GPU box the same device is listed three times (three contexts, three host threads, device-to-device copies), which
exercises the whole orchestration."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu
SUM_REL = 1e-9


def _members(nq):
    n = nq._ffi.load().nqe_device_count()
    return list(range(n)) if n >= 2 else [0, 0, 0]


def _table(nq, ctx, arrays, names):
    return nq.DeviceTable.from_arrow(pa.RecordBatch.from_arrays([pa.array(a) for a in arrays], names=names), ctx)


def _check_groups(out, keys, vals, want_avg=True):
    """out columns: key, count, sum, avg, min, max (any group order)"""
    uk, inv = np.unique(keys, return_inverse=True)
    got_key = out.column(0).to_numpy()
    o = np.argsort(got_key)
    assert np.array_equal(got_key[o], uk)
    cnt = np.bincount(inv, minlength=len(uk))
    assert out.column(1).type == pa.uint64() and np.array_equal(out.column(1).to_numpy()[o], cnt.astype(np.uint64))
    sm = np.bincount(inv, weights=vals, minlength=len(uk))
    scale = np.maximum(np.bincount(inv, weights=np.abs(vals), minlength=len(uk)), 1.0)
    assert np.all(np.abs(out.column(2).to_numpy()[o] - sm) <= SUM_REL * scale)
    assert np.all(np.abs(out.column(3).to_numpy()[o] - sm / cnt) <= SUM_REL * scale / cnt)
    mn = np.full(len(uk), np.inf); mx = np.full(len(uk), -np.inf)
    np.minimum.at(mn, inv, vals); np.maximum.at(mx, inv, vals)
    assert np.array_equal(out.column(4).to_numpy()[o], mn) and np.array_equal(out.column(5).to_numpy()[o], mx)


@pytest.mark.parametrize("dense", [True, False])
def test_multi_join_aggregate_broadcast_build(dense):
    import nqe_b200 as nq
    rng = np.random.default_rng(11 + dense)
    devs = _members(nq)
    m = nq.MultiContext(devs)
    try:
        nl, groups = 60_000, 3_000
        keys = (rng.permutation(nl) + 500).astype(np.int64) * (1 if dense else 1_000_003)  # direct table | hashed table
        a = rng.integers(0, groups, nl).astype(np.int64)
        left = _table(nq, m.members[len(devs) - 1], [keys, a], ["k", "a"])  # the build side starts on the LAST member
        shards, fks, bs = [], [], []
        for i in range(len(devs)):
            nr = 150_000 + 10_007 * i
            if i == 1:
                shards.append(None)  # a member without a shard
                continue
            fk = np.where(rng.random(nr) < 0.8, keys[rng.integers(0, nl, nr)], rng.integers(0, 400, nr)).astype(np.int64)
            b = np.round(rng.normal(0, 100, nr), 4)
            shards.append(_table(nq, m.members[i], [fk, b], ["fk", "b"]))
            fks.append(fk); bs.append(b)
        # joined schema: k, a, fk, b
        out = m.join_aggregate(left, shards, 0, 0, 1, [(5, 0), (0, 3), (1, 3), (2, 3), (3, 3), (4, 3)],
                               ["key", "count", "sum", "avg", "min", "max"]).to_arrow()
        fk, b = np.concatenate(fks), np.concatenate(bs)
        order = np.argsort(keys)
        idx = np.minimum(np.searchsorted(keys[order], fk), nl - 1)
        hit = keys[order][idx] == fk
        _check_groups(out, a[order[idx[hit]]], b[hit])
    finally:
        m.close()


def test_multi_hash_aggregate_sharded_input_and_copy():
    import nqe_b200 as nq
    rng = np.random.default_rng(5)
    devs = _members(nq)
    m = nq.MultiContext(devs)
    try:
        ks, vs, shards = [], [], []
        for i in range(len(devs)):
            n = 200_000 + 1_001 * i
            k = rng.integers(-50, 4_000, n).astype(np.int64)
            v = np.round(rng.normal(5, 50, n), 3)
            shards.append(_table(nq, m.members[i], [k, v], ["k", "v"]))
            ks.append(k); vs.append(v)
        ke, keep = nq.ColumnExpr.try_create(None, 0).to_expr(["k", "v"])
        out = m.hash_aggregate(shards, ke, [(5, 0), (0, 1), (1, 1), (2, 1), (3, 1), (4, 1)],
                               ["key", "count", "sum", "avg", "min", "max"]).to_arrow()
        _check_groups(out, np.concatenate(ks), np.concatenate(vs))
        # a table copied to another member is the same table
        moved = m.copy(shards[0], len(devs) - 1).to_arrow()
        assert np.array_equal(moved.column(0).to_numpy(), ks[0]) and np.array_equal(moved.column(1).to_numpy(), vs[0])
        # argument checks: a shard on the wrong member
        wrong = [shards[0]] * len(devs)
        if len(devs) > 1:
            with pytest.raises(nq.NqeError):
                m.hash_aggregate(wrong, ke, [(0, 1)], ["count"])
    finally:
        m.close()
