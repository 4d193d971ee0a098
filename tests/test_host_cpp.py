"""The C++ host mirror of the reference API (rustcv_b200/hostcpp/rustcv_b200.hpp) builds
against the C ABI, fails loudly without a GPU, and (GPU) matches the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_host_mirror")


def _build(oracle):
    from rustcv_b200 import build

    build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "rustcv_b200", "hostcpp"),
           "-L", os.path.join(ROOT, "rustcv_b200"), "-lrcv_imgproc", "-L", os.path.join(ROOT, "oracle"), "-lrcv_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "rustcv_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", EXE]
    subprocess.check_call(cmd)


def test_cpp_mirror_builds_and_fails_loudly_without_init(oracle):
    _build(oracle)
    out = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "ok" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_parity_on_gpu(oracle):
    _build(oracle)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_mirror ok" in out.stdout


def test_item_decode_reciprocals_are_exact(tmp_path):
    """csrc/fastdiv.h (the strip kernels decode item -> frame / band / strip with host-made reciprocals) vs real
    division, compiled as plain C++."""
    exe = str(tmp_path / "fastdiv_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "fastdiv_test.cpp"),
                           "-I", os.path.join(ROOT, "rustcv_b200", "csrc"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fastdiv ok" in out.stdout
