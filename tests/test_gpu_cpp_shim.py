"""Builds and runs the C++ program that drives the roo:: shim (include/kangaroo_b200/roo.hpp) through the
reference's own call sequence."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_runs_reference_call_sequence(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = str(tmp_path / "test_roo_shim")
    lib_dir = os.path.join(ROOT, "kangaroo_b200", "lib")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_roo_shim.cpp"), "-o", exe, "-L", lib_dir,
                           "-lroo_b200", "-Xlinker", f"-rpath={lib_dir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


def test_cpp_shim_with_the_reference_types_runs_the_application_sequence():
    """tests/cpp/test_roo_shim_kangaroo_types.cu -- the shim over the reference's own roo::Image / roo::Volume (Manage-owning)
    and the call sequence of applications/stereo2/main.cpp:375-458 -- is built where the reference headers exist
    (__graft_entry__.build()) and run here."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_roo_shim_kangaroo_types")
    if not os.path.exists(exe):
        pytest.skip("not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
