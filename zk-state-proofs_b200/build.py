"""Build the CUDA shared library (libmptv.so) in-tree with nvcc for sm_100a.

    python zk-state-proofs_b200/build.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmptv.so")
SOURCES = ["keccak_kernels.cu", "verify_kernels.cu", "rebuild_kernels.cu", "dedup_kernels.cu", "microbench.cu", "mptv_api.cu", "rebuild_api.cu", "host_codec.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-pthread", "-shared", "-cudart", "static",
    "-Xlinker", "-z,defs",  # an unresolved symbol fails the build here, not at dlopen time on the GPU box
]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "mptv.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, extra_flags=()) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", OUT] + sources()
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
