"""Builds and runs the C++ mirror of the reference's Physics interface (include/ox_b200.hpp) against the C ABI."""
import os
import subprocess

from oxide_control_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_compiles_and_runs(tmp_path):
    exe = str(tmp_path / "test_cpp_api")
    libdir = os.path.dirname(A.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "native", "test_cpp_api.cpp"), "-o", exe,
                           "-L" + libdir, "-lox_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "cpp api ok" in out.stdout


import pytest


@pytest.mark.gpu
def test_cpp_host_mirror_steps_on_the_gpu(tmp_path):
    """Same binary on the B200 box: the C++ Physics mirror must take its 'gpu path' branch (100 pendulum steps)."""
    exe = str(tmp_path / "test_cpp_api")
    libdir = os.path.dirname(A.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "native", "test_cpp_api.cpp"), "-o", exe,
                           "-L" + libdir, "-lox_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "gpu path ok" in out.stdout and "cpp api ok" in out.stdout
