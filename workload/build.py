"""Build workload/libmptgen.so (host C++ synthetic workload generator)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libmptgen.so")
SRC = os.path.join(HERE, "mptgen.cpp")


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", OUT, SRC])
    return OUT


if __name__ == "__main__":
    print(build(True))
