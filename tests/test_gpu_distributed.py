"""Two GPUs, NCCL, real peer stores: every multi-GPU plan of distributed.py against the single-process oracle.
Skipped on a one-GPU box (the CPU/gloo suite covers the orchestration there)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_plans_match_the_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import oracle as O
    from tests import dist_worker as W
    port = 29600 + os.getpid() % 2000
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    got = [np.load(str(tmp_path / f"rank{r}.npz")) for r in range(2)]
    lk, la, fk, rb = W.tables(0, W.N_BUILD, 0, W.N_PROBE)
    L = O.Batch(["k", "a"], [O.Col("i64", lk), O.Col("i64", la)])
    R = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("f64", rb)])
    joined = O.hash_join_c(L, R, 0, 0)
    want = O.aggregate(joined, ("col", 1), [("min", 1), ("count", 3), ("sum", 3), ("min", 3), ("max", 3)])

    def check_groups(m, want):
        wk = want.cols[0].values.astype(np.int64)
        wo, go = np.argsort(wk), np.argsort(m[0])
        assert np.array_equal(wk[wo], m[0][go]) and np.array_equal(want.cols[1].values[wo].astype(np.int64), m[1][go])
        assert np.allclose(want.cols[2].values[wo], m[2][go].view(np.float64), rtol=1e-9, atol=0)
        assert np.array_equal(want.cols[3].values[wo], m[3][go].view(np.float64))
        assert np.array_equal(want.cols[4].values[wo], m[4][go].view(np.float64))

    for r in range(2):
        for plan in ("broadcast", "nccl", "peer"):
            check_groups(got[r][plan], want)
            assert int(got[r][plan + "_wire"]) > 0  # rows really crossed the link
        assert {"build_all_gather", "local_join_aggregate", "partial_exchange", "merge"} <= set(got[r]["broadcast_phases"])
        T = O.Batch(["k", "v"], [O.Col("i64", (fk % 301).astype(np.int64)), O.Col("f64", rb)])
        check_groups(got[r]["group_by"], O.aggregate(T, ("col", 0), [("min", 0), ("count", 1), ("sum", 1), ("min", 1), ("max", 1)]))
    Rb = O.Batch(["fk", "b"], [O.Col("i64", fk), O.Col("i64", rb.view(np.int64).copy())])
    wj = np.stack([c.values.astype(np.int64) for c in O.hash_join_c(L, Rb, 0, 0).cols])
    assert np.array_equal(np.concatenate([g["join_broadcast"] for g in got], axis=1), wj)  # rank order = probe-row-major
    for plan in ("join_shuffle", "join_shuffle_peer"):
        js = np.concatenate([g[plan] for g in got], axis=1)
        assert js.shape == wj.shape and np.array_equal(js[:, np.lexsort(js[::-1])], wj[:, np.lexsort(wj[::-1])])
