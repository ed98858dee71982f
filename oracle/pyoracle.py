"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two checker libraries built by oracle/Makefile:

* ``Oracle``  -> oracle/_build/libmpt_oracle.so, the C restatement of
  crypto_ops::verify_merkle_proof (/root/reference/crypto-ops/src/lib.rs:8-23) and
  digest_keccak (/root/reference/crypto-ops/src/keccak.rs:6-12).
* ``RefElf``  -> oracle/_ref/librv32emu.so running the reference's own guest binary
  /root/reference/circuits/elf/riscv32im-succinct-zkvm-elf.
  Only usable where /root/reference is mounted (this container, not the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ELF = "/root/reference/circuits/elf/riscv32im-succinct-zkvm-elf"

STATUS_NAMES = {
    0: "OK",
    1: "INVALID_STATE_ROOT",
    2: "ROOT_NOT_CANONICAL",
    3: "INVALID_PROOF",
    4: "KEY_NOT_FOUND",
    5: "PANIC_OTHER",
    6: "BAD_ROOT_LEN",
    7: "DEPENDENCY_FAILED",
}


def build(force: bool = False) -> None:
    """Compile the checker libraries (gcc only; no reference source is compiled)."""
    targets = ["_build/libmpt_oracle.so", "_ref/librv32emu.so"]
    if force or not all(os.path.exists(os.path.join(HERE, t)) for t in targets):
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))
    else:
        subprocess.check_call(["make", "-s", "-C", HERE])


class _Res(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("value_node", ctypes.c_int32),
        ("value_off", ctypes.c_uint32),
        ("value_len", ctypes.c_uint32),
        ("n_perm_alg", ctypes.c_uint32),
        ("n_perm_done", ctypes.c_uint32),
    ]


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


class Oracle:
    """C restatement (oracle/mpt_oracle.c, oracle/trie_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libmpt_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.mpto_keccak256.restype = ctypes.c_uint32
        L.mpto_keccak256.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
        L.mpto_verify.restype = ctypes.c_int
        L.mpto_verify.argtypes = [
            ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
            ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32,
            ctypes.c_int, ctypes.POINTER(_Res)]
        L.mpto_account_storage_root.restype = ctypes.c_int
        L.mpto_account_storage_root.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p]
        L.mpto_verify_batch.restype = ctypes.c_int
        L.mpto_keccak256_batch.restype = ctypes.c_int
        L.mpto_trie_roots.restype = ctypes.c_int
        L.mpto_trie_get_proof.restype = ctypes.c_int

    def keccak256(self, data: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.mpto_keccak256(bytes(data), len(data), out)
        return out.raw

    def verify(self, root: bytes, proof, key: bytes, mirror: bool = False):
        """-> (status, value_bytes or None, value_node, value_off)"""
        n = len(proof)
        arr = (ctypes.c_char_p * max(n, 1))()
        lens = (ctypes.c_uint32 * max(n, 1))()
        keep = []
        for i, nd in enumerate(proof):
            b = bytes(nd)
            keep.append(b)
            arr[i] = b
            lens[i] = len(b)
        r = _Res()
        self.lib.mpto_verify(bytes(root), len(root), arr, lens, n, bytes(key), len(key),
                             1 if mirror else 0, ctypes.byref(r))
        val = None
        if r.status == 0:
            if r.value_node >= 0:
                val = keep[r.value_node][r.value_off:r.value_off + r.value_len]
            else:
                val = b""
        return r.status, val, r.value_node, r.value_off

    def account_storage_root(self, value: bytes):
        out = ctypes.create_string_buffer(32)
        rc = self.lib.mpto_account_storage_root(bytes(value), len(value), out)
        return out.raw if rc == 0 else None

    def verify_batch(self, b, nthreads: int = 1, mirror: bool = False):
        """b: dict of numpy arrays in the CSR layout of include/mptv.h.
        -> (status u8[n], value_off u64[n], value_len u32[n], perm_alg, perm_done)"""
        n = len(b["proof_first"]) - 1
        status = np.zeros(n, np.uint8)
        voff = np.zeros(n, np.uint64)
        vlen = np.zeros(n, np.uint32)
        pa = ctypes.c_uint64(0)
        pd = ctypes.c_uint64(0)
        rfp = b.get("root_from_proof")
        self.lib.mpto_verify_batch(
            _p(b["node_bytes"], ctypes.c_uint8), _p(b["node_off"], ctypes.c_uint64),
            _p(b["node_len"], ctypes.c_uint32), _p(b["proof_first"], ctypes.c_uint32),
            ctypes.c_uint64(n), _p(b["roots"], ctypes.c_uint8), _p(b["key_bytes"], ctypes.c_uint8),
            _p(b["key_off"], ctypes.c_uint32),
            _p(rfp, ctypes.c_int32) if rfp is not None else None,
            _p(status, ctypes.c_uint8), _p(voff, ctypes.c_uint64), _p(vlen, ctypes.c_uint32),
            ctypes.c_int(nthreads), ctypes.c_int(1 if mirror else 0),
            ctypes.byref(pa), ctypes.byref(pd))
        return status, voff, vlen, pa.value, pd.value

    def trie_roots(self, kv, nthreads: int = 1):
        """kv: dict(key_bytes u8, key_off u32[n+1], value_bytes u8, value_off u64[n], value_len u32[n],
        trie_first u32[T+1]) -> (roots u8[T,32], perms, nodes_hashed)"""
        T = len(kv["trie_first"]) - 1
        roots = np.zeros((T, 32), np.uint8)
        pa, nh = ctypes.c_uint64(0), ctypes.c_uint64(0)
        self.lib.mpto_trie_roots(
            _p(kv["key_bytes"], ctypes.c_uint8), _p(kv["key_off"], ctypes.c_uint32),
            _p(kv["value_bytes"], ctypes.c_uint8), _p(kv["value_off"], ctypes.c_uint64),
            _p(kv["value_len"], ctypes.c_uint32), _p(kv["trie_first"], ctypes.c_uint32), ctypes.c_uint64(T),
            _p(roots, ctypes.c_uint8), ctypes.c_int(nthreads), ctypes.byref(pa), ctypes.byref(nh))
        return roots, pa.value, nh.value

    def trie_get_proof(self, kv, t: int, key: bytes):
        """-> (root bytes, [node bytes]) for trie t of kv"""
        first = int(kv["trie_first"][t]); n = int(kv["trie_first"][t + 1]) - first
        cap = 1 << 22
        out = np.zeros(cap, np.uint8)
        lens = np.zeros(256, np.uint32)
        root = np.zeros(32, np.uint8)
        cnt = self.lib.mpto_trie_get_proof(
            _p(kv["key_bytes"], ctypes.c_uint8), _p(kv["key_off"], ctypes.c_uint32),
            _p(kv["value_bytes"], ctypes.c_uint8), _p(kv["value_off"], ctypes.c_uint64),
            _p(kv["value_len"], ctypes.c_uint32), ctypes.c_uint64(first), ctypes.c_uint64(n), bytes(key),
            ctypes.c_uint32(len(key)), _p(out, ctypes.c_uint8), ctypes.c_uint64(cap), _p(lens, ctypes.c_uint32),
            ctypes.c_uint32(256), _p(root, ctypes.c_uint8))
        nodes, pos = [], 0
        for i in range(cnt):
            nodes.append(out[pos:pos + int(lens[i])].tobytes()); pos += int(lens[i])
        return root.tobytes(), nodes

    def keccak256_batch(self, node_bytes, node_off, node_len):
        n = len(node_len)
        out = np.zeros(n * 32, np.uint8)
        self.lib.mpto_keccak256_batch(
            _p(node_bytes, ctypes.c_uint8), _p(node_off, ctypes.c_uint64),
            _p(node_len, ctypes.c_uint32), ctypes.c_uint64(n), _p(out, ctypes.c_uint8))
        return out.reshape(n, 32)


class _EmuRes(ctypes.Structure):
    _fields_ = [
        ("exit_code", ctypes.c_int32),
        ("steps", ctypes.c_uint64),
        ("pub_len", ctypes.c_uint32),
        ("err_len", ctypes.c_uint32),
        ("fault_pc", ctypes.c_uint32),
        ("fault_insn", ctypes.c_uint32),
    ]


def ref_available() -> bool:
    return os.path.exists(REF_ELF)


def classify_panic(stderr_text: str) -> int:
    """map the reference's panic text to a status class."""
    if "Invalid merkle proof" in stderr_text:
        return 1
    if "assertion" in stderr_text and "left == right" in stderr_text:
        return 2
    if "Failed to verify Merkle Proof" in stderr_text:
        return 3
    if "Key does not exist" in stderr_text:
        return 4
    return 5


def encode_hint(proof, root: bytes, key: bytes) -> bytes:
    """u64_le(len(p)) || p, p = borsh(MerkleProofInput) (crypto-ops/src/types.rs:4-9)."""
    p = struct.pack("<I", len(proof)) + b"".join(struct.pack("<I", len(n)) + bytes(n) for n in proof)
    p += struct.pack("<I", len(root)) + bytes(root) + struct.pack("<I", len(key)) + bytes(key)
    return struct.pack("<Q", len(p)) + p


class RefElf:
    """Runs the reference's own verify_merkle_proof (the committed SP1 guest ELF)."""

    CAP = 1 << 20

    def __init__(self, elf_path: str = REF_ELF):
        path = os.path.join(HERE, "_ref", "librv32emu.so")
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.lib.rv32_load.restype = ctypes.c_void_p
        self.lib.rv32_load.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
        self.lib.rv32_free.argtypes = [ctypes.c_void_p]
        self.lib.rv32_run.restype = ctypes.c_int
        self.lib.rv32_run.argtypes = [
            ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32,
            ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.POINTER(_EmuRes)]
        with open(elf_path, "rb") as f:
            self.elf = f.read()
        self._tls = threading.local()

    def _vm(self):
        vm = getattr(self._tls, "vm", None)
        if vm is None:
            vm = self.lib.rv32_load(self.elf, len(self.elf))
            if not vm:
                raise RuntimeError("rv32_load failed")
            self._tls.vm = vm
            self._tls.pub = ctypes.create_string_buffer(self.CAP)
            self._tls.err = ctypes.create_string_buffer(1 << 16)
        return vm

    def run(self, root: bytes, proof, key: bytes, max_steps: int = 0):
        """-> dict(status, value, exit_code, steps, stderr)"""
        vm = self._vm()
        hint = encode_hint(proof, root, key)
        r = _EmuRes()
        self.lib.rv32_run(vm, hint, len(hint), self._tls.pub, self.CAP, self._tls.err, 1 << 16,
                          max_steps, ctypes.byref(r))
        if r.exit_code < 0:
            raise RuntimeError(f"emulator fault pc={r.fault_pc:#x} insn={r.fault_insn:#x} code={r.exit_code}")
        err = self._tls.err.raw[:r.err_len].decode(errors="replace")
        if r.exit_code == 0:
            status, value = 0, self._tls.pub.raw[:r.pub_len]
        else:
            status, value = classify_panic(err), None
            if len(root) != 32 and "TryFromSliceError" in err or ("called `Result::unwrap()`" in err and len(root) != 32):
                status = 6
        return dict(status=status, value=value, exit_code=r.exit_code, steps=r.steps, stderr=err)
