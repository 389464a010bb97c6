"""The kernels behind the tuning knobs stay correct: the knobs are read once per process, so each alternative runs a
slice of the parity suite in a child process with the knob set.  Covers the interpreter kernels (what runs when NVRTC
is unavailable), the join table with row numbers instead of payloads, the stable split with the streaming gather,
gathered columns inside the compaction kernel, the direct (single-pass) fused join -> group-by and group-by table
paths behind the paged shared-memory ones, and the paged group-by on small inputs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ({"NQE_JIT": "0"}, "filter_project_random or kleene or null_predicate or every_operator or error_behaviour or deep_expression"),
    ({"NQE_JIT_IMPL": "ca"}, "filter_project_many_tiles and not nullable"),
    ({"NQE_JOIN_ROWPAY": "0"}, "partitioned_probe_large or join_aggregate_group_key or join_aggregate_fused"),
    ({"NQE_JOIN_SPLIT": "3", "NQE_JOIN_GATHER": "1", "NQE_JOIN_OVERLAP": "0"}, "partitioned_probe_large"),
    ({"NQE_JOIN_FUSE": "1"}, "partitioned_probe"),
    ({"NQE_JOINAGG_PAGED": "0"}, "join_aggregate_paged_large"),
    ({"NQE_AGG_PART": "0"}, "group_by_one_value_column"),
    ({"NQE_AGG_PART_MIN_ROWS": "1000"}, "group_by"),
    ({"NQE_AGG_DENSE": "0"}, "group_by_one_value or group_by_dense or paged or group_key"),
    ({"NQE_JOIN_PART_MIN_ROWS": "1000", "NQE_JOIN_PART_MIN_MB": "0"}, "hash_join or join_aggregate or golden_readme"),
    ({"NQE_JA_PROBE_SHAPE": "1", "NQE_PS_SPLIT_SHAPE": "1"}, "paged or group_by_one_value or group_key"),
    # direct join table (dense unique build keys): on every small join of the suite (duplicates, sparse keys and payload
    # collisions fall back to the hashed table); switched off, the dense-key inputs go through the hashed / paged paths
    ({"NQE_JOIN_DIRECT_MIN_ROWS": "1"}, "hash_join or join_aggregate or golden_readme or multi_batch or run_sql"),
    ({"NQE_JOIN_DIRECT": "0"}, "direct_table"),
    ({"NQE_JOIN_DIRECT_NARROW": "0"}, "direct_table"),   # 8-byte slots
    ({"NQE_JOIN_DIRECT_STAGED": "0"}, "hash_join_direct_table"),  # the general probe kernel over the direct table
]


@pytest.mark.parametrize("env,select", CASES, ids=["+".join(f"{k}={v}" for k, v in e.items()) for e, _ in CASES])
def test_alternate_kernels_pass_the_parity_slice(env, select):
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-p", "no:cacheprovider", "-k", select],
                       cwd=ROOT, env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
    tail = (p.stdout + p.stderr)[-2000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and "failed" not in p.stdout, tail
