"""Code generation of the shape-specialised filter/project kernel, checked WITHOUT a GPU: the generated CUDA source of the
three skeleton forms (plain, NULL-aware, gathered columns) must compile with NVRTC for sm_100a and must stay within the
register budget that lets two CTAs of 384 threads share an SM (65536 / 768 = 85 registers) without spilling -- the
difference between 0.40 ms and 0.57 ms per 1e8 rows (profiles/README_r01.md).  No compute call is made."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "naive-query-engine_b200")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libnqe_b200.so")):
        import __graft_entry__ as g
        g.build()
    exe = str(tmp_path_factory.mktemp("jit") / "jit_shapes")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "jit_shapes.cpp"), "-o", exe,
                           "-L", PKG, "-lnqe_b200", f"-Wl,-rpath,{PKG}"])
    return exe


@pytest.mark.parametrize("shape", ["plain", "nulls", "gather"])
def test_generated_kernel_compiles_within_the_register_budget(harness, shape, tmp_path):
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    prefix = str(tmp_path / shape)
    p = subprocess.run([harness, shape], capture_output=True, text=True, env={**os.environ, "NQE_JIT_DUMP": prefix}, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    if not os.path.exists(prefix + ".cubin"):
        pytest.skip("NVRTC not loadable here: source generated, nothing compiled")
    src = open(prefix + ".cu").read()
    assert "#define NULLS %d" % (shape == "nulls") in src and "#define GATHERS %d" % (shape == "gather") in src
    assert "#define K 8" in src  # 2048-row tiles: both rings still leave room for two CTAs per SM
    usage = subprocess.run([CUOBJDUMP, "-res-usage", prefix + ".cubin"], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+)", usage)
    assert m, usage
    regs, stack = int(m.group(1)), int(m.group(2))
    assert regs <= 80 and stack == 0, usage
    sass = subprocess.run([CUOBJDUMP, "-sass", prefix + ".cubin"], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass  # cp.async.bulk staging into the shared-memory rings
    assert not re.search(r"\b(STL|LDL)\b", sass)  # no register spills
