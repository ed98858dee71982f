"""Build the CUDA shared library (libmptv.so) in-tree with nvcc for sm_100a.

    python zk-state-proofs_b200/build.py [--force]

Every source is compiled to its own object (in parallel, rebuilt only when it or a header changed) under
build/obj/, then linked into zk-state-proofs_b200/libmptv.so.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmptv.so")
OBJ = os.path.join(HERE, "..", "build", "obj")
SOURCES = ["keccak_kernels.cu", "verify_kernels.cu", "single_kernels.cu", "borsh_kernels.cu", "rebuild_kernels.cu", "dedup_kernels.cu",
           "microbench.cu", "mptv_api.cu", "rebuild_api.cu", "host_codec.cpp", "host_flatten.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-pthread,-mavx2"]
LINK_FLAGS = ARCH + ["-shared", "-cudart", "static", "-Xcompiler", "-pthread",
                     "-Xlinker", "-z,defs"]  # an unresolved symbol fails the build here, not at dlopen time on the GPU box


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))]
    return hs + [os.path.join(HERE, "..", "include", "mptv.h"), os.path.abspath(__file__)]


def _obj_of(src, tag):
    return os.path.join(OBJ, tag + os.path.basename(src) + ".o")


def build(force: bool = False, extra_flags=(), out: str = OUT) -> str:
    """extra_flags (e.g. -DMPTV_WALK_MINB=3) build a variant: its objects are cached under their own tag."""
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    tag = ("v" + str(abs(hash(tuple(extra_flags))) % 10**8) + "_") if extra_flags else ""
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    todo = []
    for s in sources():
        o = _obj_of(s, tag)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            todo.append((s, o))

    def cc(so):
        s, o = so
        subprocess.check_call([nvcc] + COMPILE_FLAGS + list(extra_flags) + ["-c", s, "-o", o])

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(cc, todo))
    objs = [_obj_of(s, tag) for s in sources()]
    if todo or not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc] + LINK_FLAGS + ["-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
