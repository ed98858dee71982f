"""Builds and runs tests/cpp/test_crypto_ops.cpp: the C++ reference-shaped host API
(include/mptv_crypto_ops.hpp) over libmptv.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_crypto_ops")


def build_exe():
    import zk_state_proofs_b200 as z
    z.load_library()
    src = EXE + ".cpp"
    deps = [src, os.path.join(ROOT, "include", "mptv_crypto_ops.hpp"), os.path.join(ROOT, "include", "mptv.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        libdir = os.path.dirname(z.lib_path())
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", libdir, "-l:libmptv.so", "-Wl,-rpath," + libdir, "-pthread", "-ldl"])
    return EXE


def test_cpp_host_codecs_cpu():
    import torch
    exe = build_exe()
    mode = "--cpu" if torch.cuda.is_available() else "--cpu-nogpu"
    r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_shaped_api_gpu():
    r = subprocess.run([build_exe(), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
