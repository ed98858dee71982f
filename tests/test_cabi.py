"""The C-ABI shared library loads and exports every symbol include/mptv.h declares (no compute
calls: there is no GPU here), and the host mirror's wire types round-trip."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mptv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mptv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import zk_state_proofs_b200 as z
    from importlib import import_module
    if not os.path.exists(z.lib_path()):
        import subprocess, sys
        subprocess.check_call([sys.executable, os.path.join(ROOT, "zk-state-proofs_b200", "build.py")])
    lib = ctypes.CDLL(z.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s


def test_strerror_and_status_names_without_a_gpu():
    import zk_state_proofs_b200 as z
    L = z.load_library()
    assert b"no CPU fallback" in L.mptv_strerror(-6)
    assert L.mptv_status_name(4) == b"KEY_NOT_FOUND"
    for k, v in z.STATUS_NAMES.items():
        assert L.mptv_status_name(k).decode() == v


def test_no_gpu_means_loud_failure():
    import torch
    import zk_state_proofs_b200 as z
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(z.MptvError):
        z.Verifier([0])
    with pytest.raises(z.MptvError):
        z.digest_keccak(b"abc")


def test_borsh_roundtrip_and_flatten_layout():
    import numpy as np
    import zk_state_proofs_b200 as z
    inp = z.MerkleProofInput([b"\x01" * 5, b"", b"\x02" * 33], b"\xaa" * 32, b"\x12\x34")
    assert z.MerkleProofInput.from_borsh(inp.to_borsh()) == inp
    b = z.flatten([inp, z.MerkleProofInput([b"\x07" * 17], b"\xbb" * 31, b"")])
    assert b.n_proofs == 2 and b.n_nodes == 4
    assert (b.node_off % 16 == 0).all()
    assert list(b.node_len) == [5, 0, 33, 17]
    assert list(b.proof_first) == [0, 3, 4]
    assert b.bad_root_len is not None and list(b.bad_root_len) == [False, True]
    assert b.node_bytes[int(b.node_off[2]):int(b.node_off[2]) + 33].tobytes() == b"\x02" * 33
    assert b.n_perm() == 4


def test_wire_forms_of_both_input_types_roundtrip():
    """borsh (what the prover feeds the guests, prover/src/bin/main.rs:41) and serde JSON (types.rs derives both)"""
    import json
    import zk_state_proofs_b200 as z
    m = z.MerkleProofInput([b"\x01\x02", b"", b"\xff" * 40], b"\xaa" * 32, b"\x80")
    assert z.MerkleProofInput.from_json(m.to_json()) == m
    assert json.loads(m.to_json())["key"] == [128]
    s = z.StorageProofInput([b"\x01" * 3, b"\x02" * 70], [[b"\x03" * 5], [], [b"\x04", b"\x05" * 33]], b"\xbb" * 32,
                            b"\x11" * 20, [b"\x21" * 32, b"", b"\x22" * 32], bytes(range(32)))
    w = s.to_borsh()
    assert z.StorageProofInput.from_borsh(w) == s and z.StorageProofInput.from_json(s.to_json()) == s
    # borsh layout: u32 count, then u32-prefixed byte vectors; the fixed [u8; 32] has no length prefix
    assert w[:4] == (2).to_bytes(4, "little") and w[4:8] == (3).to_bytes(4, "little") and w[-32:] == bytes(range(32))
    with pytest.raises(ValueError):
        z.StorageProofInput.from_borsh(w + b"\x00")


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the boundary is a C ABI: include/mptv.h compiles as C99 (-pedantic) and a C program links libmptv.so"""
    import subprocess
    import zk_state_proofs_b200 as z
    z.load_library()
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(z.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "cabi_smoke.c"), "-o", exe, "-L", libdir, "-l:libmptv.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "rlp(300) has 3 bytes" in out.stdout and "storage entries: ok" in out.stdout


@pytest.mark.gpu
def test_c_program_runs_the_storage_guest_entry_on_the_gpu(tmp_path):
    """the same plain-C program on a B200: mptv_verify_storage_borsh answers from C"""
    import subprocess
    import zk_state_proofs_b200 as z
    z.load_library()
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(z.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "cabi_smoke.c"), "-o", exe, "-L", libdir, "-l:libmptv.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "mptv_verify_storage_borsh -> 0, input status 1" in out.stdout and "storage entries: ok" in out.stdout, out.stdout + out.stderr


def test_binding_structures_match_the_pinned_abi_layout():
    """include/mptv.h pins sizeof / offsetof of every structure (MPTV_ABI_PIN: the header does not compile if one
    moves); the ctypes mirrors must have the same sizes and field offsets"""
    import ctypes
    from zk_state_proofs_b200 import crypto_ops as co
    sizes = dict(_CBatch=88, _CResult=24, _CKvBatch=72, _CProofTargets=32, _CProofsOut=64, Timings=64, RebuildTimings=64,
                 HostStats=120, FlattenInfo=32, _CLog=40)
    for name, want in sizes.items():
        assert ctypes.sizeof(getattr(co, name)) == want, name
    assert [getattr(co._CBatch, f).offset for f, _ in co._CBatch._fields_] == [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80]
    assert [getattr(co._CResult, f).offset for f, _ in co._CResult._fields_] == [0, 8, 16]
    assert co._CKvBatch.n_items.offset == 48 and co._CKvBatch.n_tries.offset == 64 and co._CProofsOut.n_nodes.offset == 48
    src = open(os.path.join(ROOT, "include", "mptv.h")).read()
    assert src.count("MPTV_ABI_PIN(") >= 25
