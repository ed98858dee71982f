"""Host-side mirror of the reference's crypto-ops crate over the C ABI (include/mptv.h).

Reference interface mirrored here (same names, argument meaning and error behaviour):
  verify_merkle_proof(root_hash, proof, key) -> bytes     /root/reference/crypto-ops/src/lib.rs:8-23
      panics in the reference -> raises VerifyPanic carrying the same message class
  digest_keccak(data) -> 32 bytes                         /root/reference/crypto-ops/src/keccak.rs:6-12
  MerkleProofInput / StorageProofInput (+ borsh wire form) /root/reference/crypto-ops/src/types.rs:4-19
and the batched entry the north star adds:
  verify_merkle_proofs([MerkleProofInput]) -> [bytes | VerifyPanic]
The tx / receipt trie rebuild that trie-utils performs with eth_trie
(/root/reference/trie-utils/src/proofs/transaction.rs:41-66, proofs/receipt.rs:49-84) is
trie_roots() / ordered_trie_root() over a KvBatch.
The nested account -> storage flow of the risc0 storage guest
(/root/reference/circuits/risc0-storage-proof/storage-proof-circuit/storage-circuit/src/main.rs:6-31)
is verify_storage_proof_input().

Everything here is plumbing: flatten to the CSR arena, call libmptv.so, slice the results.  No
hashing, RLP decoding or trie walking happens on the CPU, and nothing under oracle/ is imported.
"""
from __future__ import annotations

import ctypes
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

STATUS_NAMES = {
    0: "OK", 1: "INVALID_STATE_ROOT", 2: "ROOT_NOT_CANONICAL", 3: "INVALID_PROOF",
    4: "KEY_NOT_FOUND", 5: "PANIC_OTHER", 6: "BAD_ROOT_LEN", 7: "DEPENDENCY_FAILED",
}
# the reference's panic messages per verdict class (crypto-ops/src/lib.rs:14,19,21,22)
PANIC_MESSAGES = {
    1: "Invalid merkle proof",
    2: "assertion `left == right` failed",
    3: "Failed to verify Merkle Proof: InvalidProof",
    4: "Key does not exist!",
    5: "panicked inside eth_trie (invalid data / index out of bounds)",
    6: "called `Result::unwrap()` on an `Err` value: TryFromSliceError",
    7: "account proof rejected or not an Account RLP",
}


class MptvError(RuntimeError):
    """libmptv.so missing, no B200, or a negative MPTV_ERR_* from the C ABI."""


class VerifyPanic(Exception):
    """The reference would have panicked; .status is the MPTV_ST_* class."""

    def __init__(self, status: int):
        self.status = status
        super().__init__(f"{PANIC_MESSAGES.get(status, '?')} [{STATUS_NAMES.get(status, status)}]")

    def __eq__(self, other):
        return isinstance(other, VerifyPanic) and other.status == self.status

    def __hash__(self):
        return hash(("VerifyPanic", self.status))


# ----------------------------------------------------------------------------- wire types
def _borsh_bytes(b: bytes) -> bytes:
    return struct.pack("<I", len(b)) + bytes(b)


def _read_vec_u8(buf: memoryview, pos: int):
    (n,) = struct.unpack_from("<I", buf, pos)
    pos += 4
    return bytes(buf[pos:pos + n]), pos + n


@dataclass
class MerkleProofInput:
    """crypto-ops/src/types.rs:4-9"""
    proof: List[bytes]
    root_hash: bytes
    key: bytes

    def to_borsh(self) -> bytes:
        out = struct.pack("<I", len(self.proof)) + b"".join(_borsh_bytes(n) for n in self.proof)
        return out + _borsh_bytes(self.root_hash) + _borsh_bytes(self.key)

    @classmethod
    def from_borsh(cls, data: bytes) -> "MerkleProofInput":
        buf = memoryview(data)
        (n,) = struct.unpack_from("<I", buf, 0)
        pos = 4
        proof = []
        for _ in range(n):
            b, pos = _read_vec_u8(buf, pos)
            proof.append(b)
        root, pos = _read_vec_u8(buf, pos)
        key, pos = _read_vec_u8(buf, pos)
        if pos != len(data):
            raise ValueError("trailing bytes after MerkleProofInput")
        return cls(proof, root, key)

    # serde form (types.rs:4 derives Serialize / Deserialize): byte vectors are JSON arrays of numbers
    def to_json(self) -> str:
        import json
        return json.dumps(dict(proof=[list(n) for n in self.proof], root_hash=list(self.root_hash), key=list(self.key)))

    @classmethod
    def from_json(cls, text: str) -> "MerkleProofInput":
        import json
        d = json.loads(text)
        return cls([bytes(n) for n in d["proof"]], bytes(d["root_hash"]), bytes(d["key"]))


@dataclass
class StorageProofInput:
    """crypto-ops/src/types.rs:11-19 (storage keys are un-hashed; the consumer hashes them)"""
    account_proof: List[bytes]
    storage_proofs: List[List[bytes]]
    root_hash: bytes
    account_key: bytes
    storage_keys: List[bytes]
    address_keccak: bytes

    def to_borsh(self) -> bytes:
        out = struct.pack("<I", len(self.account_proof)) + b"".join(_borsh_bytes(n) for n in self.account_proof)
        out += struct.pack("<I", len(self.storage_proofs))
        for pr in self.storage_proofs:
            out += struct.pack("<I", len(pr)) + b"".join(_borsh_bytes(n) for n in pr)
        out += _borsh_bytes(self.root_hash) + _borsh_bytes(self.account_key)
        out += struct.pack("<I", len(self.storage_keys)) + b"".join(_borsh_bytes(k) for k in self.storage_keys)
        return out + bytes(self.address_keccak)

    @classmethod
    def from_borsh(cls, data: bytes) -> "StorageProofInput":
        buf = memoryview(data)

        def vec_vec(pos):
            (n,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            out = []
            for _ in range(n):
                b, pos = _read_vec_u8(buf, pos)
                out.append(b)
            return out, pos

        account_proof, pos = vec_vec(0)
        (ns,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        storage_proofs = []
        for _ in range(ns):
            pr, pos = vec_vec(pos)
            storage_proofs.append(pr)
        root, pos = _read_vec_u8(buf, pos)
        akey, pos = _read_vec_u8(buf, pos)
        skeys, pos = vec_vec(pos)
        if len(data) - pos != 32:
            raise ValueError("StorageProofInput must end with the 32-byte address_keccak")
        return cls(account_proof, storage_proofs, root, akey, skeys, bytes(buf[pos:pos + 32]))

    def to_json(self) -> str:
        import json
        return json.dumps(dict(account_proof=[list(n) for n in self.account_proof],
                               storage_proofs=[[list(n) for n in pr] for pr in self.storage_proofs],
                               root_hash=list(self.root_hash), account_key=list(self.account_key),
                               storage_keys=[list(k) for k in self.storage_keys],
                               address_keccak=list(self.address_keccak)))

    @classmethod
    def from_json(cls, text: str) -> "StorageProofInput":
        import json
        d = json.loads(text)
        return cls([bytes(n) for n in d["account_proof"]], [[bytes(n) for n in pr] for pr in d["storage_proofs"]],
                   bytes(d["root_hash"]), bytes(d["account_key"]), [bytes(k) for k in d["storage_keys"]],
                   bytes(d["address_keccak"]))


# ----------------------------------------------------------------------------- CSR batch
@dataclass
class Batch:
    """Flat CSR form of n MerkleProofInputs (layout: include/mptv.h `mptv_batch`)."""
    node_bytes: np.ndarray   # u8, nodes 16-byte aligned, total padded to 16
    node_off: np.ndarray     # u64 [n_nodes]
    node_len: np.ndarray     # u32 [n_nodes]
    proof_first: np.ndarray  # u32 [n_proofs+1]
    roots: np.ndarray        # u8 [32*n_proofs]
    key_bytes: np.ndarray    # u8
    key_off: np.ndarray      # u32 [n_proofs+1]
    root_from_proof: Optional[np.ndarray] = None  # i32 [n_proofs]
    bad_root_len: Optional[np.ndarray] = None     # bool [n_proofs]: root_hash.len() != 32

    @property
    def n_proofs(self) -> int:
        return len(self.proof_first) - 1

    @property
    def n_nodes(self) -> int:
        return len(self.node_len)

    def n_perm(self) -> int:
        """Algorithmic Keccak-f count: sum over nodes of ceil((len+1)/136)."""
        return int((self.node_len.astype(np.int64) // 136 + 1).sum())

    def value(self, off: int, ln: int) -> bytes:
        return self.node_bytes[off:off + ln].tobytes()


def flatten(inputs: Sequence[MerkleProofInput], root_from_proof: Optional[Sequence[int]] = None) -> Batch:
    """MerkleProofInput x n  ->  CSR arena (every node on a 16-byte boundary)."""
    n = len(inputs)
    lens = np.fromiter((len(nd) for inp in inputs for nd in inp.proof), dtype=np.uint32)
    counts = np.fromiter((len(inp.proof) for inp in inputs), dtype=np.int64, count=n)
    proof_first = np.zeros(n + 1, np.uint32)
    np.cumsum(counts, out=proof_first[1:])
    padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
    node_off = np.zeros(len(lens), np.uint64)
    if len(lens):
        np.cumsum(padded[:-1], out=node_off[1:])
    total = int(padded.sum())
    node_bytes = np.zeros(total + 16, np.uint8)
    i = 0
    for inp in inputs:
        for nd in inp.proof:
            o = int(node_off[i])
            node_bytes[o:o + len(nd)] = np.frombuffer(bytes(nd), np.uint8)
            i += 1
    roots = np.zeros(32 * n, np.uint8)
    bad = np.zeros(n, bool)
    klens = np.fromiter((len(inp.key) for inp in inputs), dtype=np.int64, count=n)
    key_off = np.zeros(n + 1, np.uint32)
    np.cumsum(klens, out=key_off[1:])
    key_bytes = np.zeros(int(key_off[-1]) + 16, np.uint8)
    for p, inp in enumerate(inputs):
        if len(inp.root_hash) == 32:
            roots[32 * p:32 * p + 32] = np.frombuffer(bytes(inp.root_hash), np.uint8)
        else:
            bad[p] = True
        if inp.key:
            key_bytes[int(key_off[p]):int(key_off[p + 1])] = np.frombuffer(bytes(inp.key), np.uint8)
    rfp = None
    if root_from_proof is not None:
        rfp = np.asarray(root_from_proof, np.int32)
    return Batch(node_bytes, node_off, lens, proof_first, roots, key_bytes, key_off, rfp, bad if bad.any() else None)


class _HostBatchOwner:
    """keeps an mptv_host_batch (C++-owned arrays) alive while numpy views of it exist"""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def __del__(self):
        try:
            if self.handle:
                self.lib.mptv_host_batch_free(self.handle)
                self.handle = None
        except Exception:
            pass


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    ct = {np.uint8: ctypes.c_uint8, np.uint32: ctypes.c_uint32, np.uint64: ctypes.c_uint64}[dtype]
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(int(n),))


def flatten_borsh(blobs: Sequence[bytes], threads: int = 0, pinned: bool = False) -> Batch:
    """borsh(MerkleProofInput) x n -> CSR batch through the multi-threaded C++ flattener
    (csrc/host_codec.cpp, mptv_flatten_borsh).  The arrays are views of C++-owned (optionally
    page-locked) memory kept alive by the returned Batch."""
    L = load_library()
    n = len(blobs)
    lens = np.fromiter((len(x) for x in blobs), dtype=np.int64, count=n)
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(bytes(x) for x in blobs) + b"\0", np.uint8)
    h = ctypes.c_void_p()
    rc = L.mptv_flatten_borsh(buf.ctypes.data, off.ctypes.data, n, threads, 1 if pinned else 0, ctypes.byref(h))
    if rc != 0:
        raise ValueError(f"mptv_flatten_borsh: {L.mptv_strerror(rc).decode()} (malformed borsh MerkleProofInput?)")
    b = batch_from_handle(L, h, n)
    b._owner = _HostBatchOwner(L, h)
    return b


class FlattenInfo(ctypes.Structure):
    """include/mptv.h `mptv_flatten_info`"""
    _fields_ = [("n_nodes", ctypes.c_uint64), ("nodes_aliased", ctypes.c_uint64),
                ("node_bytes_supplied", ctypes.c_uint64), ("node_bytes_placed", ctypes.c_uint64)]


class HostStats(ctypes.Structure):
    """include/mptv.h `mptv_host_stats`"""
    _fields_ = [("chunks", ctypes.c_uint64), ("nodes", ctypes.c_uint64), ("nodes_aliased", ctypes.c_uint64),
                ("node_bytes_supplied", ctypes.c_uint64), ("node_bytes_placed", ctypes.c_uint64),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("flatten_us", ctypes.c_uint64), ("wait_us", ctypes.c_uint64), ("map_us", ctypes.c_uint64),
                ("call_us", ctypes.c_uint64), ("launches", ctypes.c_uint64), ("pull_chunks", ctypes.c_uint64),
                ("device_chunks", ctypes.c_uint64), ("index_us", ctypes.c_uint64)]


FLATTEN_ALIAS_DUPLICATES = 1


def _blob_arrays(blobs, blob_off=None):
    """sequence of bytes objects, or one uint8 array + [n + 1] uint64 offsets -> (uint8 array, uint64 offsets)"""
    if blob_off is None:
        n = len(blobs)
        lens = np.fromiter((len(x) for x in blobs), dtype=np.int64, count=n)
        blob_off = np.zeros(n + 1, np.uint64)
        np.cumsum(lens, out=blob_off[1:])
        blobs = np.frombuffer(b"".join(bytes(x) for x in blobs) + b"\0", np.uint8)
    return np.ascontiguousarray(blobs, np.uint8), np.ascontiguousarray(blob_off, np.uint64)


def flatten_borsh_ex(blobs, blob_off=None, threads: int = 0, pinned: bool = False, alias_duplicates: bool = True):
    """mptv_flatten_borsh_ex: one pass over the blobs; byte-identical nodes are stored once and aliased.
    -> (Batch, FlattenInfo).  The Batch's node_bytes block also holds the index arrays."""
    L = load_library()
    buf, off = _blob_arrays(blobs, blob_off)
    n = len(off) - 1
    h = ctypes.c_void_p()
    info = FlattenInfo()
    rc = L.mptv_flatten_borsh_ex(buf.ctypes.data, off.ctypes.data, n, threads, 1 if pinned else 0,
                                 FLATTEN_ALIAS_DUPLICATES if alias_duplicates else 0, ctypes.byref(h), ctypes.byref(info))
    if rc != 0:
        raise ValueError(f"mptv_flatten_borsh_ex: {L.mptv_strerror(rc).decode()} (malformed borsh MerkleProofInput?)")
    b = batch_from_handle(L, h, n)
    b._owner = _HostBatchOwner(L, h)
    return b, info


def flatten_storage_borsh(blobs, blob_off=None, threads: int = 0, pinned: bool = False, alias_duplicates: bool = False):
    """mptv_flatten_storage_borsh: borsh(StorageProofInput) blobs -> (Batch with root_from_proof, hash_key u8[n_proofs],
    proof_first u64[n_inputs + 1], FlattenInfo): what Verifier.verify_batch_hashed_keys takes."""
    L = load_library()
    buf, off = _blob_arrays(blobs, blob_off)
    n = len(off) - 1
    h = ctypes.c_void_p()
    info = FlattenInfo()
    pf = np.zeros(n + 1, np.uint64)
    hk = ctypes.c_void_p()
    L.mptv_flatten_storage_borsh.restype = ctypes.c_int32
    L.mptv_flatten_storage_borsh.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rc = L.mptv_flatten_storage_borsh(buf.ctypes.data, off.ctypes.data, n, threads, 1 if pinned else 0,
                                      FLATTEN_ALIAS_DUPLICATES if alias_duplicates else 0, ctypes.byref(h), ctypes.byref(info),
                                      pf.ctypes.data, ctypes.byref(hk))
    if rc != 0:
        raise ValueError(f"mptv_flatten_storage_borsh: {L.mptv_strerror(rc).decode()} (malformed borsh StorageProofInput?)")
    npr = int(pf[n])
    b = batch_from_handle(L, h, npr)
    v = ctypes.cast(L.mptv_host_batch_view(h), ctypes.POINTER(_CBatch)).contents
    b.root_from_proof = _view(v.root_from_proof, npr, np.uint32).view(np.int32)
    b._owner = _HostBatchOwner(L, h)
    return b, _view(hk.value, npr, np.uint8).copy(), pf, info


def account_storage_root(value: bytes) -> Optional[bytes]:
    """mptv_account_storage_root: the storage guest's alloy_rlp::decode_exact::<Account>(value) (main.rs:15) ->
    the 32-byte storage_root, or None where the guest's unwrap() panics."""
    L = load_library()
    L.mptv_account_storage_root.restype = ctypes.c_int
    L.mptv_account_storage_root.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p]
    out = ctypes.create_string_buffer(32)
    return out.raw if L.mptv_account_storage_root(bytes(value), len(value), out) else None


def borsh_flatten_probe(blobs, blob_off=None, threads: int = 0, chunk_bytes: int = 32 << 20, alias_duplicates: bool = True):
    """mptv_borsh_flatten_probe: the host stage of verify_borsh alone -> (seconds, FlattenInfo)"""
    import time
    L = load_library()
    buf, off = _blob_arrays(blobs, blob_off)
    info = FlattenInfo()
    t0 = time.perf_counter()
    rc = L.mptv_borsh_flatten_probe(buf.ctypes.data, off.ctypes.data, len(off) - 1, threads, chunk_bytes,
                                    1 if alias_duplicates else 0, ctypes.byref(info))
    dt = time.perf_counter() - t0
    if rc != 0:
        raise ValueError(f"mptv_borsh_flatten_probe: {L.mptv_strerror(rc).decode()}")
    return dt, info


def host_bw_probe(threads: int = 0, bytes_per_thread: int = 128 << 20):
    """mptv_host_bw_probe -> (read GB/s, copy GB/s of payload) of the host's memory system with `threads` threads"""
    L = load_library()
    r, c = ctypes.c_double(), ctypes.c_double()
    rc = L.mptv_host_bw_probe(threads, bytes_per_thread, ctypes.byref(r), ctypes.byref(c))
    if rc != 0:
        raise MptvError(f"mptv_host_bw_probe: {L.mptv_strerror(rc).decode()}")
    return r.value, c.value


def batch_from_handle(L, h, n: int) -> Batch:
    """Batch view of a mptv_host_batch handle (no ownership: the caller frees or recycles the handle)."""
    v = ctypes.cast(L.mptv_host_batch_view(h), ctypes.POINTER(_CBatch)).contents
    bad = _view(L.mptv_host_batch_bad_root(h), n, np.uint8).astype(bool)
    return Batch(_view(v.node_bytes, v.node_bytes_len, np.uint8), _view(v.node_off, v.n_nodes, np.uint64),
                 _view(v.node_len, v.n_nodes, np.uint32), _view(v.proof_first, n + 1, np.uint32),
                 _view(v.roots, 32 * n, np.uint8), _view(v.key_bytes, int(_view(v.key_off, n + 1, np.uint32)[-1]) + 16, np.uint8),
                 _view(v.key_off, n + 1, np.uint32), None, bad if bad.any() else None)


@dataclass
class Log:
    """trie-utils/src/types.rs:11-15"""
    address: bytes            # 20 bytes
    topics: List[bytes]       # 32 bytes each
    data: bytes


class _CLog(ctypes.Structure):
    _fields_ = [("address", ctypes.c_char_p), ("topics", ctypes.c_char_p), ("n_topics", ctypes.c_uint32),
                ("data", ctypes.c_char_p), ("data_len", ctypes.c_uint32)]


def encode_receipt(status: bool, cumulative_gas_used: int, bloom: bytes, logs: Sequence[Log],
                   prefix: Optional[int] = None) -> bytes:
    """The leaf bytes insert_receipt puts into the receipt trie (trie-utils/src/receipt.rs:8-38):
    [prefix] ++ rlp([status, cumulative_gas_used, bloom, logs]); prefix None = legacy receipt."""
    L = load_library()
    if len(bloom) != 256 or any(len(l.address) != 20 or any(len(t) != 32 for t in l.topics) for l in logs):
        raise ValueError("bloom must be 256 bytes, addresses 20, topics 32")
    arr = (_CLog * max(1, len(logs)))()
    keep = []
    for i, l in enumerate(logs):
        t = b"".join(l.topics)
        keep.append((bytes(l.address), t, bytes(l.data)))
        arr[i] = _CLog(keep[-1][0], keep[-1][1], len(l.topics), keep[-1][2], len(l.data))
    p = -1 if prefix is None else int(prefix)
    n = L.mptv_encode_receipt(p, 1 if status else 0, int(cumulative_gas_used), bytes(bloom), arr, len(logs), None, 0)
    out = ctypes.create_string_buffer(n)
    L.mptv_encode_receipt(p, 1 if status else 0, int(cumulative_gas_used), bytes(bloom), arr, len(logs), out, n)
    return out.raw


# ----------------------------------------------------------------------------- key/value batches (rebuild)
@dataclass
class KvBatch:
    """Flat form of T independent tries' (key, value) insert lists (include/mptv.h `mptv_kv_batch`)."""
    key_bytes: np.ndarray    # u8
    key_off: np.ndarray      # u32 [n_items+1]
    value_bytes: np.ndarray  # u8, values 16-byte aligned, total padded to 16
    value_off: np.ndarray    # u64 [n_items]
    value_len: np.ndarray    # u32 [n_items]; 0 = delete (eth_trie: insert(k, b"") removes k)
    trie_first: np.ndarray   # u32 [n_tries+1]

    @property
    def n_items(self) -> int:
        return len(self.value_len)

    @property
    def n_tries(self) -> int:
        return len(self.trie_first) - 1

    def as_dict(self):
        return dict(key_bytes=self.key_bytes, key_off=self.key_off, value_bytes=self.value_bytes,
                    value_off=self.value_off, value_len=self.value_len, trie_first=self.trie_first)


def rlp_index(i: int) -> bytes:
    """alloy_rlp::encode(index) -- the trie key of transaction / receipt i (transaction.rs:45)."""
    if i == 0:
        return b"\x80"
    if i < 0x80:
        return bytes([i])
    be = i.to_bytes((i.bit_length() + 7) // 8, "big")
    return bytes([0x80 + len(be)]) + be


def rlp_index_native(i: int) -> bytes:
    """the same through the C++ host codec (mptv_rlp_index)"""
    out = ctypes.create_string_buffer(9)
    n = load_library().mptv_rlp_index(int(i), out)
    return out.raw[:n]


def flatten_kv(tries) -> KvBatch:
    """[[(key, value), ...], ...] in insertion order -> KvBatch (every value on a 16-byte boundary)."""
    keys = [k for t in tries for k, _ in t]
    vals = [v for t in tries for _, v in t]
    n = len(keys)
    klens = np.fromiter((len(k) for k in keys), dtype=np.int64, count=n)
    key_off = np.zeros(n + 1, np.uint32)
    np.cumsum(klens, out=key_off[1:])
    key_bytes = np.zeros(int(key_off[-1]) + 16, np.uint8)
    if n:
        key_bytes[:int(key_off[-1])] = np.frombuffer(b"".join(bytes(k) for k in keys), np.uint8)
    value_len = np.fromiter((len(v) for v in vals), dtype=np.uint32, count=n)
    padded = (value_len.astype(np.uint64) + 15) & ~np.uint64(15)
    value_off = np.zeros(n, np.uint64)
    if n:
        np.cumsum(padded[:-1], out=value_off[1:])
    value_bytes = np.zeros(int(padded.sum()) + 16, np.uint8)
    for v, o in zip(vals, value_off):
        if len(v):
            value_bytes[int(o):int(o) + len(v)] = np.frombuffer(bytes(v), np.uint8)
    counts = np.fromiter((len(t) for t in tries), dtype=np.int64, count=len(tries))
    trie_first = np.zeros(len(tries) + 1, np.uint32)
    np.cumsum(counts, out=trie_first[1:])
    return KvBatch(key_bytes, key_off, value_bytes, value_off, value_len, trie_first)


# ----------------------------------------------------------------------------- C ABI
class _CKvBatch(ctypes.Structure):
    _fields_ = [
        ("key_bytes", ctypes.c_void_p), ("key_off", ctypes.c_void_p), ("value_bytes", ctypes.c_void_p),
        ("value_bytes_len", ctypes.c_uint64), ("value_off", ctypes.c_void_p), ("value_len", ctypes.c_void_p),
        ("n_items", ctypes.c_uint64), ("trie_first", ctypes.c_void_p), ("n_tries", ctypes.c_uint64),
    ]


class _CProofTargets(ctypes.Structure):
    _fields_ = [("trie", ctypes.c_void_p), ("key_bytes", ctypes.c_void_p), ("key_off", ctypes.c_void_p),
                ("n_targets", ctypes.c_uint64)]


class _CProofsOut(ctypes.Structure):
    _fields_ = [("node_bytes", ctypes.c_void_p), ("node_bytes_cap", ctypes.c_uint64), ("node_off", ctypes.c_void_p),
                ("node_len", ctypes.c_void_p), ("nodes_cap", ctypes.c_uint64), ("proof_first", ctypes.c_void_p),
                ("n_nodes", ctypes.c_uint64), ("node_bytes_len", ctypes.c_uint64)]


class RebuildTimings(ctypes.Structure):
    _fields_ = [
        ("structure_ms", ctypes.c_float), ("encode_ms", ctypes.c_float), ("keccak_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float), ("n_nodes", ctypes.c_uint64), ("n_hashed", ctypes.c_uint64),
        ("n_perm", ctypes.c_uint64), ("arena_bytes", ctypes.c_uint64), ("levels", ctypes.c_uint32),
        ("keccak_launches", ctypes.c_uint32), ("other_launches", ctypes.c_uint32), ("pad", ctypes.c_uint32),
    ]


class _CBatch(ctypes.Structure):
    _fields_ = [
        ("node_bytes", ctypes.c_void_p), ("node_bytes_len", ctypes.c_uint64),
        ("node_off", ctypes.c_void_p), ("node_len", ctypes.c_void_p), ("n_nodes", ctypes.c_uint64),
        ("proof_first", ctypes.c_void_p), ("n_proofs", ctypes.c_uint64), ("roots", ctypes.c_void_p),
        ("key_bytes", ctypes.c_void_p), ("key_off", ctypes.c_void_p), ("root_from_proof", ctypes.c_void_p),
    ]


class _CResult(ctypes.Structure):
    _fields_ = [("status", ctypes.c_void_p), ("value_off", ctypes.c_void_p), ("value_len", ctypes.c_void_p)]


class Timings(ctypes.Structure):
    _fields_ = [
        ("bin_ms", ctypes.c_float), ("keccak_ms", ctypes.c_float), ("parse_ms", ctypes.c_float),
        ("walk_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("n_nodes", ctypes.c_uint64),
        ("n_perm", ctypes.c_uint64), ("keccak_launches", ctypes.c_uint32), ("other_launches", ctypes.c_uint32),
        ("n_unique_nodes", ctypes.c_uint64), ("n_unique_perm", ctypes.c_uint64),
    ]


def lib_path() -> str:
    return os.environ.get("MPTV_LIB", os.path.join(HERE, "libmptv.so"))


_LIB = None


def load_library():
    """dlopen libmptv.so (built in-tree by build.py).  Raises MptvError when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MptvError(f"{path} not found: build it with `python zk-state-proofs_b200/build.py` "
                        "(there is no CPU fallback)")
    L = ctypes.CDLL(path)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    L.mptv_create.restype = i32
    L.mptv_create.argtypes = [ctypes.POINTER(ctypes.c_int), i32, ctypes.POINTER(vp)]
    L.mptv_destroy.restype = None
    L.mptv_destroy.argtypes = [vp]
    L.mptv_device_count.restype = i32
    L.mptv_device_count.argtypes = [vp]
    L.mptv_last_error.restype = ctypes.c_char_p
    L.mptv_last_error.argtypes = [vp]
    L.mptv_strerror.restype = ctypes.c_char_p
    L.mptv_strerror.argtypes = [i32]
    L.mptv_status_name.restype = ctypes.c_char_p
    L.mptv_status_name.argtypes = [i32]
    L.mptv_verify_batch.restype = i32
    L.mptv_verify_batch.argtypes = [vp, ctypes.POINTER(_CBatch), ctypes.POINTER(_CResult)]
    L.mptv_verify_borsh.restype = i32
    L.mptv_verify_borsh.argtypes = [vp, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(_CResult)]
    L.mptv_verify_storage_borsh.restype = i32
    L.mptv_verify_storage_borsh.argtypes = [vp, vp, vp, ctypes.c_uint64, ctypes.c_int, vp, vp, ctypes.c_uint64, ctypes.POINTER(_CResult)]
    L.mptv_verify_batch_hashed_keys.restype = i32
    L.mptv_verify_batch_hashed_keys.argtypes = [vp, ctypes.POINTER(_CBatch), vp, ctypes.POINTER(_CResult)]
    L.mptv_verify_batch_device.restype = i32
    L.mptv_verify_batch_device.argtypes = [vp, i32, ctypes.POINTER(_CBatch), ctypes.POINTER(_CResult), vp]
    L.mptv_keccak256_batch.restype = i32
    L.mptv_keccak256_batch.argtypes = [vp, vp, u64, vp, vp, u64, vp]
    L.mptv_keccak256_batch_device.restype = i32
    L.mptv_keccak256_batch_device.argtypes = [vp, i32, vp, vp, vp, u64, vp, vp]
    L.mptv_last_timings.restype = i32
    L.mptv_last_timings.argtypes = [vp, i32, ctypes.POINTER(Timings)]
    L.mptv_int_issue_peak.restype = i32
    L.mptv_int_issue_peak.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.mptv_set_option.restype = i32
    L.mptv_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.mptv_trie_roots.restype = i32
    L.mptv_trie_roots.argtypes = [vp, ctypes.POINTER(_CKvBatch), vp]
    L.mptv_trie_roots_device.restype = i32
    L.mptv_trie_roots_device.argtypes = [vp, i32, ctypes.POINTER(_CKvBatch), vp, vp]
    L.mptv_last_rebuild_timings.restype = i32
    L.mptv_last_rebuild_timings.argtypes = [vp, i32, ctypes.POINTER(RebuildTimings)]
    L.mptv_trie_proofs.restype = i32
    L.mptv_trie_proofs.argtypes = [vp, ctypes.POINTER(_CKvBatch), ctypes.POINTER(_CProofTargets), vp,
                                   ctypes.POINTER(_CProofsOut)]
    L.mptv_flatten_borsh.restype = i32
    L.mptv_flatten_borsh.argtypes = [vp, vp, u64, i32, i32, ctypes.POINTER(vp)]
    L.mptv_flatten_borsh_ex.restype = i32
    L.mptv_flatten_borsh_ex.argtypes = [vp, vp, u64, i32, i32, ctypes.c_uint, ctypes.POINTER(vp), ctypes.POINTER(FlattenInfo)]
    L.mptv_borsh_flatten_probe.restype = i32
    L.mptv_borsh_flatten_probe.argtypes = [vp, vp, u64, i32, u64, i32, ctypes.POINTER(FlattenInfo)]
    L.mptv_host_bw_probe.restype = i32
    L.mptv_host_bw_probe.argtypes = [i32, u64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.mptv_host_stats_get.restype = i32
    L.mptv_host_stats_get.argtypes = [vp, ctypes.POINTER(HostStats), i32]
    L.mptv_host_batch_view.restype = vp
    L.mptv_host_batch_view.argtypes = [vp]
    L.mptv_host_batch_bad_root.restype = vp
    L.mptv_host_batch_bad_root.argtypes = [vp]
    L.mptv_host_batch_free.restype = None
    L.mptv_host_batch_free.argtypes = [vp]
    L.mptv_rlp_index.restype = ctypes.c_uint32
    L.mptv_rlp_index.argtypes = [u64, ctypes.c_char_p]
    L.mptv_encode_receipt.restype = u64
    L.mptv_encode_receipt.argtypes = [i32, i32, u64, ctypes.c_char_p, ctypes.POINTER(_CLog), ctypes.c_uint32,
                                      ctypes.c_char_p, u64]
    L.mptv_alloc_pinned.restype = vp
    L.mptv_alloc_pinned.argtypes = [ctypes.c_size_t]
    L.mptv_free_pinned.restype = None
    L.mptv_free_pinned.argtypes = [vp]
    _LIB = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Verifier:
    """Owns an mptv_ctx (device memory + streams on the chosen GPUs)."""

    def __init__(self, device_ids: Optional[Sequence[int]] = None):
        self.lib = load_library()
        self.ctx = ctypes.c_void_p()
        if device_ids is None:
            rc = self.lib.mptv_create(None, 0, ctypes.byref(self.ctx))
        else:
            arr = (ctypes.c_int * len(device_ids))(*device_ids)
            rc = self.lib.mptv_create(arr, len(device_ids), ctypes.byref(self.ctx))
        if rc != 0:
            raise MptvError(f"mptv_create failed: {self.lib.mptv_strerror(rc).decode()}")

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.mptv_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise MptvError(f"{what}: {self.lib.mptv_strerror(rc).decode()} "
                            f"({self.lib.mptv_last_error(self.ctx).decode()})")

    @property
    def device_count(self) -> int:
        return self.lib.mptv_device_count(self.ctx)

    def set_option(self, name: str, value: int):
        self._check(self.lib.mptv_set_option(self.ctx, name.encode(), int(value)), f"set_option({name})")

    # -- host-buffer entry (H2D + kernels + D2H inside the call)
    def verify_batch(self, b: Batch):
        n = b.n_proofs
        status = np.zeros(n, np.uint8)
        voff = np.zeros(n, np.uint64)
        vlen = np.zeros(n, np.uint32)
        if n == 0:
            return status, voff, vlen
        cb = _CBatch(_ptr(b.node_bytes), len(b.node_bytes), _ptr(b.node_off), _ptr(b.node_len), b.n_nodes,
                     _ptr(b.proof_first), n, _ptr(b.roots), _ptr(b.key_bytes), _ptr(b.key_off),
                     _ptr(b.root_from_proof))
        cr = _CResult(_ptr(status), _ptr(voff), _ptr(vlen))
        self._check(self.lib.mptv_verify_batch(self.ctx, ctypes.byref(cb), ctypes.byref(cr)), "mptv_verify_batch")
        if b.bad_root_len is not None:
            status[b.bad_root_len] = 6
            voff[b.bad_root_len] = 0
            vlen[b.bad_root_len] = 0
        return status, voff, vlen

    def verify_batch_hashed_keys(self, b: Batch, hash_key: np.ndarray):
        """mptv_verify_batch_hashed_keys: proofs flagged in hash_key are looked up under keccak256(key)."""
        n = b.n_proofs
        status = np.zeros(n, np.uint8)
        voff = np.zeros(n, np.uint64)
        vlen = np.zeros(n, np.uint32)
        if n == 0:
            return status, voff, vlen
        hk = np.ascontiguousarray(hash_key, np.uint8)
        cb = _CBatch(_ptr(b.node_bytes), len(b.node_bytes), _ptr(b.node_off), _ptr(b.node_len), b.n_nodes,
                     _ptr(b.proof_first), n, _ptr(b.roots), _ptr(b.key_bytes), _ptr(b.key_off),
                     _ptr(b.root_from_proof))
        cr = _CResult(_ptr(status), _ptr(voff), _ptr(vlen))
        self._check(self.lib.mptv_verify_batch_hashed_keys(self.ctx, ctypes.byref(cb), _ptr(hk), ctypes.byref(cr)),
                    "mptv_verify_batch_hashed_keys")
        if b.bad_root_len is not None:
            status[b.bad_root_len] = 6
            voff[b.bad_root_len] = 0
            vlen[b.bad_root_len] = 0
        return status, voff, vlen

    def verify_borsh(self, blobs, blob_off=None, threads: int = 0):
        """mptv_verify_borsh: borsh(MerkleProofInput) blobs in, verdicts out, flattening pipelined with the copies
        and kernels.  `blobs` is a sequence of bytes objects, or one uint8 array with `blob_off` ([n + 1] uint64).
        Returns (status, value_off, value_len); value_off indexes the CONCATENATED blobs."""
        buf, off = _blob_arrays(blobs, blob_off)
        n = len(off) - 1
        status = np.zeros(n, np.uint8)
        voff = np.zeros(n, np.uint64)
        vlen = np.zeros(n, np.uint32)
        if n <= 0:
            return status, voff, vlen
        cr = _CResult(_ptr(status), _ptr(voff), _ptr(vlen))
        self._check(self.lib.mptv_verify_borsh(self.ctx, buf.ctypes.data, off.ctypes.data, n, threads, ctypes.byref(cr)),
                    "mptv_verify_borsh")
        return status, voff, vlen

    def verify_storage_borsh(self, blobs, blob_off=None, threads: int = 0, n_proofs: Optional[int] = None):
        """mptv_verify_storage_borsh: borsh(StorageProofInput) blobs in (the storage guest's own input,
        storage-circuit/src/main.rs:6-9), the guest's flow for every input out.  Returns
        (proof_first u64[n + 1], input_status u8[n], status, value_off, value_len): input i owns the results
        [proof_first[i], proof_first[i + 1]) -- its account proof, then its storage proofs; value_off indexes the
        CONCATENATED blobs.  n_proofs: the total, when the caller knows it (saves the sizing call)."""
        buf, off = _blob_arrays(blobs, blob_off)
        n = len(off) - 1
        pf = np.zeros(max(n, 0) + 1, np.uint64)
        ist = np.zeros(max(n, 0), np.uint8)
        if n <= 0:
            return pf, ist, np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        args = (self.ctx, buf.ctypes.data, off.ctypes.data, n, threads, _ptr(pf), _ptr(ist))
        if n_proofs is None:
            rc = self.lib.mptv_verify_storage_borsh(*args, 0, None)  # first pass only: how many proofs the guest verifies
            if rc != -4:  # MPTV_ERR_NOMEM: proof_first now holds the layout
                self._check(rc, "mptv_verify_storage_borsh")
            n_proofs = int(pf[n])
        for attempt in range(2):
            status = np.zeros(n_proofs, np.uint8)
            voff = np.zeros(n_proofs, np.uint64)
            vlen = np.zeros(n_proofs, np.uint32)
            cr = _CResult(_ptr(status), _ptr(voff), _ptr(vlen))
            rc = self.lib.mptv_verify_storage_borsh(*args, n_proofs, ctypes.byref(cr))
            if rc == -4 and attempt == 0 and int(pf[n]) > n_proofs:  # the caller's count was too small: proof_first[n] is the real one
                n_proofs = int(pf[n])
                continue
            self._check(rc, "mptv_verify_storage_borsh")
            break
        return pf, ist, status[:int(pf[n])], voff[:int(pf[n])], vlen[:int(pf[n])]

    def verify_storage_proof_inputs_borsh(self, inputs: Sequence[StorageProofInput]):
        """verify_storage_proof_inputs through the wire format: the inputs are serialised as the prover would write them
        and streamed through mptv_verify_storage_borsh.  -> per input: the list of storage values, or the VerifyPanic
        the guest would have died with."""
        blobs = [inp.to_borsh() for inp in inputs]
        buf, off = _blob_arrays(blobs, None)
        pf, ist, status, voff, vlen = self.verify_storage_borsh(buf, off)
        out = []
        for i in range(len(blobs)):
            if ist[i]:
                out.append(VerifyPanic(int(ist[i])))
            else:
                out.append([buf[int(voff[q]):int(voff[q]) + int(vlen[q])].tobytes() for q in range(int(pf[i]) + 1, int(pf[i + 1]))])
        return out

    # -- device-resident entry: pointers are raw device addresses (ints), e.g. torch tensors' data_ptr()
    def verify_batch_device(self, dev_index: int, ptrs: dict, n_nodes: int, n_proofs: int, out_ptrs: dict,
                            stream: int = 0, node_bytes_len: int = 0):
        cb = _CBatch(ptrs["node_bytes"], node_bytes_len, ptrs["node_off"], ptrs["node_len"], n_nodes,
                     ptrs["proof_first"], n_proofs, ptrs["roots"], ptrs["key_bytes"], ptrs["key_off"],
                     ptrs.get("root_from_proof"))
        cr = _CResult(out_ptrs["status"], out_ptrs["value_off"], out_ptrs["value_len"])
        self._check(self.lib.mptv_verify_batch_device(self.ctx, dev_index, ctypes.byref(cb), ctypes.byref(cr),
                                                      ctypes.c_void_p(stream) if stream else None),
                    "mptv_verify_batch_device")

    def keccak256_batch_device(self, dev_index: int, node_bytes: int, node_off: int, node_len: int, n_nodes: int,
                               digests: int, stream: int = 0):
        self._check(self.lib.mptv_keccak256_batch_device(self.ctx, dev_index, node_bytes, node_off, node_len,
                                                         n_nodes, digests,
                                                         ctypes.c_void_p(stream) if stream else None),
                    "mptv_keccak256_batch_device")

    def host_stats(self, reset: bool = False) -> HostStats:
        """what the host-fed entries moved (chunks, nodes aliased, bytes over PCIe) since the last reset"""
        t = HostStats()
        self._check(self.lib.mptv_host_stats_get(self.ctx, ctypes.byref(t), 1 if reset else 0), "mptv_host_stats_get")
        return t

    def last_timings(self, dev_index: int = 0) -> Timings:
        t = Timings()
        self._check(self.lib.mptv_last_timings(self.ctx, dev_index, ctypes.byref(t)), "mptv_last_timings")
        return t

    def int_issue_peak(self, dev_index: int = 0, mode: int = 0) -> float:
        """32-bit lane-operations / s of the alu pipe (0 = LOP3, 1 = SHF, 2 = Keccak mix)."""
        v = ctypes.c_double()
        self._check(self.lib.mptv_int_issue_peak(self.ctx, dev_index, mode, ctypes.byref(v)), "mptv_int_issue_peak")
        return v.value

    def keccak256_batch(self, node_bytes: np.ndarray, node_off: np.ndarray, node_len: np.ndarray) -> np.ndarray:
        n = len(node_len)
        out = np.zeros((n, 32), np.uint8)
        if n:
            self._check(self.lib.mptv_keccak256_batch(self.ctx, _ptr(node_bytes), len(node_bytes), _ptr(node_off),
                                                      _ptr(node_len), n, _ptr(out)), "mptv_keccak256_batch")
        return out

    # -- trie rebuild (EthTrie::new / insert x n / root_hash for a batch of tries)
    def trie_roots(self, kv: KvBatch) -> np.ndarray:
        """-> u8 [n_tries, 32]; host buffers, sharded over the context's devices."""
        roots = np.zeros((kv.n_tries, 32), np.uint8)
        if kv.n_tries == 0:
            return roots
        cb = _CKvBatch(_ptr(kv.key_bytes), _ptr(kv.key_off), _ptr(kv.value_bytes), len(kv.value_bytes),
                       _ptr(kv.value_off), _ptr(kv.value_len), kv.n_items, _ptr(kv.trie_first), kv.n_tries)
        self._check(self.lib.mptv_trie_roots(self.ctx, ctypes.byref(cb), _ptr(roots)), "mptv_trie_roots")
        return roots

    def trie_roots_device(self, dev_index: int, ptrs: dict, n_items: int, n_tries: int, roots_ptr: int,
                          stream: int = 0, value_bytes_len: int = 0):
        cb = _CKvBatch(ptrs["key_bytes"], ptrs["key_off"], ptrs["value_bytes"], value_bytes_len, ptrs["value_off"],
                       ptrs["value_len"], n_items, ptrs["trie_first"], n_tries)
        self._check(self.lib.mptv_trie_roots_device(self.ctx, dev_index, ctypes.byref(cb), roots_ptr,
                                                    ctypes.c_void_p(stream) if stream else None),
                    "mptv_trie_roots_device")

    def last_rebuild_timings(self, dev_index: int = 0) -> RebuildTimings:
        t = RebuildTimings()
        self._check(self.lib.mptv_last_rebuild_timings(self.ctx, dev_index, ctypes.byref(t)),
                    "mptv_last_rebuild_timings")
        return t

    def trie_proofs(self, kv: KvBatch, targets):
        """Rebuild the tries and extract Trie::get_proof for every target (trie index, key bytes).
        -> (roots u8[n_tries, 32], Batch) where proof q of the batch belongs to target q and its root /
        key fields are already filled in, i.e. the batch can be verified as it is."""
        nq = len(targets)
        roots = np.zeros((kv.n_tries, 32), np.uint8)
        t_trie = np.fromiter((t for t, _ in targets), dtype=np.uint32, count=nq)
        klens = np.fromiter((len(k) for _, k in targets), dtype=np.int64, count=nq)
        key_off = np.zeros(nq + 1, np.uint32)
        np.cumsum(klens, out=key_off[1:])
        key_bytes = np.zeros(int(key_off[-1]) + 16, np.uint8)
        if nq:
            key_bytes[:int(key_off[-1])] = np.frombuffer(b"".join(bytes(k) for _, k in targets), np.uint8)
        cb = _CKvBatch(_ptr(kv.key_bytes), _ptr(kv.key_off), _ptr(kv.value_bytes), len(kv.value_bytes),
                       _ptr(kv.value_off), _ptr(kv.value_len), kv.n_items, _ptr(kv.trie_first), kv.n_tries)
        tg = _CProofTargets(_ptr(t_trie), _ptr(key_bytes), _ptr(key_off), nq)
        proof_first = np.zeros(nq + 1, np.uint32)
        cap_nodes, cap_bytes = 8 * nq + 8, 1024 * nq + 4096
        for _ in range(2):
            node_bytes = np.zeros(cap_bytes, np.uint8)
            node_off = np.zeros(cap_nodes, np.uint64)
            node_len = np.zeros(cap_nodes, np.uint32)
            out = _CProofsOut(_ptr(node_bytes), cap_bytes, _ptr(node_off), _ptr(node_len), cap_nodes,
                              _ptr(proof_first), 0, 0)
            rc = self.lib.mptv_trie_proofs(self.ctx, ctypes.byref(cb), ctypes.byref(tg), _ptr(roots), ctypes.byref(out))
            if rc != -4:  # MPTV_ERR_NOMEM: the struct now holds the required capacities
                break
            cap_nodes, cap_bytes = int(out.n_nodes) + 8, int(out.node_bytes_len) + 16
        self._check(rc, "mptv_trie_proofs")
        nn, nb = int(out.n_nodes), int(out.node_bytes_len)
        b = Batch(node_bytes[:nb], node_off[:nn], node_len[:nn], proof_first,
                  np.ascontiguousarray(roots[t_trie].reshape(-1)) if nq else np.zeros(0, np.uint8),
                  key_bytes, key_off, None, None)
        return roots, b

    def transaction_proof_inputs(self, encoded_txs: Sequence[bytes], target_index: int) -> MerkleProofInput:
        """The RPC-free half of get_ethereum_transaction_proof_inputs (transaction.rs:41-73) and
        get_ethereum_receipt_proof_inputs (receipt.rs:49-92): trie {rlp(i): encoded item i}, root_hash(),
        get_proof(rlp(target_index))."""
        kv = flatten_kv([[(rlp_index(i), v) for i, v in enumerate(encoded_txs)]])
        key = rlp_index(target_index)
        roots, b = self.trie_proofs(kv, [(0, key)])
        proof = [b.node_bytes[int(o):int(o) + int(n)].tobytes() for o, n in zip(b.node_off, b.node_len)]
        return MerkleProofInput(proof, roots[0].tobytes(), key)

    def ordered_trie_root(self, values: Sequence[bytes]) -> bytes:
        """Root of the trie {rlp(i): values[i]} -- what transaction.rs:41-66 / receipt.rs:49-84 compute."""
        kv = flatten_kv([[(rlp_index(i), v) for i, v in enumerate(values)]])
        return self.trie_roots(kv)[0].tobytes()

    # -- reference-shaped API
    def verify_merkle_proofs(self, inputs: Sequence[MerkleProofInput], root_from_proof=None):
        """-> list of bytes (the value) or VerifyPanic (what the reference would have panicked with)."""
        b = flatten(inputs, root_from_proof)
        status, voff, vlen = self.verify_batch(b)
        out = []
        for p in range(len(inputs)):
            if status[p] == 0:
                out.append(b.value(int(voff[p]), int(vlen[p])))
            else:
                out.append(VerifyPanic(int(status[p])))
        return out

    def verify_merkle_proof(self, root_hash: bytes, proof: Sequence[bytes], key: bytes) -> bytes:
        r = self.verify_merkle_proofs([MerkleProofInput(list(proof), bytes(root_hash), bytes(key))])[0]
        if isinstance(r, VerifyPanic):
            raise r
        return r

    def digest_keccak(self, data: bytes) -> bytes:
        nb = np.zeros(((len(data) + 15) // 16) * 16 + 16, np.uint8)
        nb[:len(data)] = np.frombuffer(bytes(data), np.uint8)
        return self.keccak256_batch(nb, np.zeros(1, np.uint64), np.array([len(data)], np.uint32))[0].tobytes()

    def verify_storage_proof_input(self, inp: StorageProofInput) -> List[bytes]:
        """The risc0 storage guest (storage-circuit/src/main.rs:6-31): account proof under
        address_keccak, then every storage proof under the account's storage_root with key
        keccak(storage_key).  Returns the verified storage values; raises VerifyPanic like the guest.
        ONE C-ABI call: the storage keys are hashed on the device (mptv_verify_batch_hashed_keys) and the roots of
        the storage proofs are taken on the device from the verified account leaf."""
        r = self.verify_storage_proof_inputs([inp])[0]
        if isinstance(r, VerifyPanic):
            raise r
        return r

    def verify_storage_proof_inputs(self, inputs: Sequence[StorageProofInput]):
        """Batched storage guest: every input's account proof and all of its storage proofs in ONE
        device batch (storage keys hashed on the device from the caller's raw keys, storage roots taken on the
        device from the verified account leaves).  -> per input: list of storage values, or the VerifyPanic the
        guest would have died with (the first failing proof in the guest's order)."""
        # the guest zips storage_proofs with storage_keys (main.rs:18-21): the shorter list decides
        items, rfp, hk, spans = [], [], [], []
        for inp in inputs:
            a = len(items)
            items.append(MerkleProofInput(inp.account_proof, inp.root_hash, bytes(inp.address_keccak)))
            rfp.append(-1)
            hk.append(0)
            for pr, key in zip(inp.storage_proofs, inp.storage_keys):
                items.append(MerkleProofInput(pr, b"\x00" * 32, bytes(key)))  # raw slot: digest_keccak(&key) happens on the device
                rfp.append(a)
                hk.append(1)
            spans.append((a, len(items)))
        res = []
        if items:
            b = flatten(items, rfp)
            status, voff, vlen = self.verify_batch_hashed_keys(b, np.array(hk, np.uint8))
            res = [b.value(int(voff[p]), int(vlen[p])) if status[p] == 0 else VerifyPanic(int(status[p])) for p in range(len(items))]
        out = []
        for a, e in spans:
            bad = next((r for r in res[a:e] if isinstance(r, VerifyPanic)), None)
            # decode_exact::<Account>(..).unwrap() (main.rs:15) runs even when no storage proof follows
            if bad is None and e == a + 1 and account_storage_root(res[a]) is None:
                bad = VerifyPanic(7)
            out.append(bad if bad is not None else res[a + 1:e])
        return out

    def _keccak_many(self, datas: Sequence[bytes]) -> List[bytes]:
        if not datas:
            return []
        lens = np.array([len(d) for d in datas], np.uint32)
        padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
        off = np.zeros(len(datas), np.uint64)
        np.cumsum(padded[:-1], out=off[1:])
        nb = np.zeros(int(padded.sum()) + 16, np.uint8)
        for d, o in zip(datas, off):
            nb[int(o):int(o) + len(d)] = np.frombuffer(bytes(d), np.uint8)
        return [r.tobytes() for r in self.keccak256_batch(nb, off, lens)]


_DEFAULT: Optional[Verifier] = None


def _default() -> Verifier:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Verifier([0])
    return _DEFAULT


def verify_merkle_proof(root_hash: bytes, proof: Sequence[bytes], key: bytes) -> bytes:
    """Drop-in for crypto_ops::verify_merkle_proof (lib.rs:8): returns the value or raises VerifyPanic."""
    return _default().verify_merkle_proof(root_hash, proof, key)


def verify_merkle_proofs(inputs: Sequence[MerkleProofInput]):
    return _default().verify_merkle_proofs(inputs)


def digest_keccak(data: bytes) -> bytes:
    """Drop-in for crypto_ops::keccak::digest_keccak (keccak.rs:6)."""
    return _default().digest_keccak(data)


def verify_storage_proof_input(inp: StorageProofInput) -> List[bytes]:
    return _default().verify_storage_proof_input(inp)


def trie_roots(kv: KvBatch) -> np.ndarray:
    return _default().trie_roots(kv)


def ordered_trie_root(values: Sequence[bytes]) -> bytes:
    return _default().ordered_trie_root(values)
