"""The C++ host mirror (naive-query-engine_b200/host/physical_plan.hpp) and its reference-style
tests (tests/cpp/test_physical_plan.cpp): compiled on CPU, run on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "naive-query-engine_b200")


def _build(tmp_path):
    if not os.path.exists(os.path.join(PKG, "libnqe_b200.so")):
        import __graft_entry__ as g
        g.build()
    exe = str(tmp_path / "test_physical_plan")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_physical_plan.cpp"), "-o", exe,
                           "-L", PKG, "-lnqe_b200", f"-Wl,-rpath,{PKG}"])
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_cuda(tmp_path):
    exe = _build(tmp_path)
    import nqe_b200 as nq
    if nq.load().nqe_device_count() > 0:
        pytest.skip("CUDA device present (covered by the gpu test)")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "CudaError" in p.stderr  # no CPU fallback


@pytest.mark.gpu
def test_cpp_reference_style_tests_on_gpu(tmp_path):
    exe = _build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "ALL OK" in p.stdout
