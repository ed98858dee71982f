"""workload/gen.py -- synthetic workloads of the shapes BASELINE.json names, as CSR batches.

Test / bench infrastructure (stands in for the RPC half of trie-utils, see mptgen.cpp).  Produces
`zk_state_proofs_b200.Batch` objects; never touches oracle/ and is never called by the product.

  config 2: n_proofs account proofs against an n_accounts state trie           -> account_batch()
  config 3: groups of 1 account + k storage-slot proofs, mixed incl/excl/mutated -> nested_batch()
  config 5: mixed account / storage proofs                                      -> mixed_batch()
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import build as _build

MUT_NONE, MUT_FLIP_LEAF, MUT_FLIP_INNER, MUT_DROP_LAST, MUT_DROP_ROOT, MUT_WRONG_KEY, MUT_SHUFFLE, MUT_JUNK = range(8)

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build.build())
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        L.mptgen_trie_build.restype = vp
        L.mptgen_trie_build.argtypes = [u64, u64, i32, vp, ctypes.c_uint32, i32]
        L.mptgen_trie_free.argtypes = [vp]
        L.mptgen_trie_info.argtypes = [vp, vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(u64)]
        L.mptgen_trie_entry.restype = ctypes.c_uint32
        L.mptgen_trie_entry.argtypes = [vp, u64, vp, vp]
        L.mptgen_proofs_plan.argtypes = [vp, vp, vp, u64, u64, vp, vp, i32]
        L.mptgen_proofs_emit.argtypes = [vp, vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        L.mptgen_trie_raw_keys.argtypes = [vp, i32]
        L.mptgen_csr_to_storage_borsh.restype = u64
        L.mptgen_csr_to_storage_borsh.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, i32]
        L.mptgen_keccak256.argtypes = [vp, u64, vp]
        L.mptgen_csr_to_borsh.restype = u64
        L.mptgen_csr_to_borsh.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, i32]
        _LIB = L
    return _LIB


def _threads() -> int:
    return max(1, min(32, os.cpu_count() or 1))


class SynthTrie:
    """A synthetic state (kind=0) or ERC-20 storage (kind=1) trie held in host memory."""

    def __init__(self, n_keys: int, seed: int, kind: int = 0, pool_roots: Optional[np.ndarray] = None):
        self.n_keys, self.seed, self.kind = n_keys, seed, kind
        pr = None if pool_roots is None else np.ascontiguousarray(pool_roots, np.uint8)
        self.h = lib().mptgen_trie_build(n_keys, seed, kind, None if pr is None else pr.ctypes.data,
                                         0 if pr is None else len(pr) // 32, _threads())
        if not self.h:
            raise MemoryError("mptgen_trie_build failed")
        root = np.zeros(32, np.uint8)
        nn, ab, nk = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        lib().mptgen_trie_info(self.h, root.ctypes.data, ctypes.byref(nn), ctypes.byref(ab), ctypes.byref(nk))
        self.root = root
        self.n_nodes, self.arena_bytes = nn.value, ab.value

    def entry(self, i: int):
        key = np.zeros(32, np.uint8)
        val = np.zeros(256, np.uint8)
        n = lib().mptgen_trie_entry(self.h, i, key.ctypes.data, val.ctypes.data)
        return key.tobytes(), val[:n].tobytes()

    def close(self):
        if self.h:
            lib().mptgen_trie_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan(self, sel: np.ndarray, mut: np.ndarray, seed2: int):
        n = len(sel)
        nc = np.zeros(n, np.uint32)
        bc = np.zeros(n, np.uint64)
        lib().mptgen_proofs_plan(self.h, sel.ctypes.data, mut.ctypes.data, n, seed2, nc.ctypes.data,
                                 bc.ctypes.data, _threads())
        return nc, bc

    def emit(self, sel, mut, seed2, slot, proof_first, byte_first, node_bytes, node_off, node_len, roots, keys32, raw32=None):
        lib().mptgen_proofs_emit(self.h, sel.ctypes.data, mut.ctypes.data, len(sel), seed2,
                                 None if slot is None else slot.ctypes.data, proof_first.ctypes.data,
                                 byte_first.ctypes.data, node_bytes.ctypes.data, node_off.ctypes.data,
                                 node_len.ctypes.data, roots.ctypes.data, keys32.ctypes.data, _threads(),
                                 None if raw32 is None else raw32.ctypes.data)

    def set_raw_keys(self, on: bool = True):
        """storage tries: every proof key gets a pre-image (an absent / wrong key is keccak(32 random bytes) instead of
        32 random bytes), so that the batch has a StorageProofInput wire form; set before plan / emit"""
        lib().mptgen_trie_raw_keys(self.h, 1 if on else 0)


def _alloc(n, dtype, pinned):
    if pinned:
        import torch
        tdt = {np.uint8: torch.uint8, np.uint32: torch.int32, np.uint64: torch.int64, np.int32: torch.int32}[dtype]
        t = torch.empty(max(int(n), 1), dtype=tdt, pin_memory=True)
        a = t.numpy().view(dtype)[:int(n)]
        a[...] = 0
        _KEEP.append(t)
        return a
    return np.zeros(int(n), dtype)


_KEEP = []  # pinned torch tensors backing numpy views


def draw_mix(rng: np.random.Generator, n: int, p_excl: float, p_mut: float, n_keys: int):
    """-> (sel i64[n] (-1 = absent key), mut u8[n])"""
    sel = rng.integers(0, n_keys, n, dtype=np.int64)
    u = rng.random(n)
    sel[u < p_excl] = -1
    mut = np.zeros(n, np.uint8)
    m = (u >= p_excl) & (u < p_excl + p_mut)
    mut[m] = rng.integers(1, 8, int(m.sum()), dtype=np.uint8)
    return sel, mut


def _assemble(parts, n_proofs, rfp, pinned, raw_keys=False):
    """parts: list of (trie, sel, mut, seed2, slot u64[]) covering every proof slot exactly once.
    raw_keys: also keep every key's pre-image (batch.raw_keys u8[32 n]: the storage slot; for an account proof its key)."""
    import zk_state_proofs_b200 as z
    nc_all = np.zeros(n_proofs, np.uint32)
    bc_all = np.zeros(n_proofs, np.uint64)
    for trie, sel, mut, seed2, slot in parts:
        nc, bc = trie.plan(sel, mut, seed2)
        nc_all[slot] = nc
        bc_all[slot] = bc
    proof_first = _alloc(n_proofs + 1, np.uint32, pinned)
    np.cumsum(nc_all, out=proof_first[1:])
    byte_first = np.zeros(n_proofs + 1, np.uint64)
    np.cumsum(bc_all, out=byte_first[1:])
    n_nodes, total = int(proof_first[-1]), int(byte_first[-1])
    node_bytes = _alloc(total + 16, np.uint8, pinned)
    node_off = _alloc(n_nodes, np.uint64, pinned)
    node_len = _alloc(n_nodes, np.uint32, pinned)
    roots = _alloc(32 * n_proofs, np.uint8, pinned)
    keys = _alloc(32 * n_proofs + 16, np.uint8, pinned)
    raw = np.zeros(32 * n_proofs, np.uint8) if raw_keys else None
    for trie, sel, mut, seed2, slot in parts:
        trie.emit(sel, mut, seed2, slot, proof_first, byte_first, node_bytes, node_off, node_len, roots, keys, raw)
    key_off = _alloc(n_proofs + 1, np.uint32, pinned)
    key_off[:] = np.arange(n_proofs + 1, dtype=np.uint32) * 32
    r = None
    if rfp is not None:
        r = _alloc(n_proofs, np.int32, pinned)
        r[:] = rfp
    b = z.Batch(node_bytes, node_off, node_len, proof_first, roots, keys, key_off, r, None)
    if raw_keys:
        b.raw_keys = raw
    return b


def account_batch(trie: SynthTrie, n_proofs: int, seed: int, p_excl: float = 0.0, p_mut: float = 0.0,
                  pinned: bool = False, force_mut: int = 0):
    """config 2: account proofs for addresses sampled uniformly from the trie.  force_mut (1 ... 7) applies that one
    mutator to EVERY proof (6 = node order shuffled: still accepted by the reference, but nothing is chain-shaped)."""
    rng = np.random.default_rng(seed)
    sel, mut = draw_mix(rng, n_proofs, p_excl, p_mut, trie.n_keys)
    if force_mut:
        mut[:] = force_mut
    slot = np.arange(n_proofs, dtype=np.uint64)
    return _assemble([(trie, sel, mut, seed, slot)], n_proofs, None, pinned)


def nested_batch(state: SynthTrie, tokens, n_groups: int, seed: int, k_storage: int = 3, p_excl: float = 0.10,
                 p_mut: float = 0.10, pinned: bool = False, raw_keys: bool = False):
    """config 3: groups of (1 account proof + k storage proofs whose root is the account's
    storage_root).  `state` must have been built with pool_roots = [t.root for t in tokens].
    raw_keys: the storage keys keep their pre-images (batch.raw_keys), see SynthTrie.set_raw_keys -> batch_to_storage_borsh."""
    for tok in tokens:
        tok.set_raw_keys(raw_keys)
    rng = np.random.default_rng(seed)
    P = len(tokens)
    G = 1 + k_storage
    n = n_groups * G
    acct = rng.integers(0, state.n_keys, n_groups, dtype=np.int64)
    a_sel, a_mut = acct.copy(), np.zeros(n_groups, np.uint8)
    u = rng.random(n_groups)
    m = u < p_mut
    a_mut[m] = rng.integers(1, 8, int(m.sum()), dtype=np.uint8)
    parts = [(state, a_sel, a_mut, seed, (np.arange(n_groups, dtype=np.uint64) * G))]
    tok_of_group = (acct % P).astype(np.int64)
    for t_idx, tok in enumerate(tokens):
        groups = np.nonzero(tok_of_group == t_idx)[0]
        if len(groups) == 0:
            continue
        cnt = len(groups) * k_storage
        sel, mut = draw_mix(rng, cnt, p_excl, p_mut, tok.n_keys)
        slot = (np.repeat(groups.astype(np.uint64) * G, k_storage) +
                np.tile(np.arange(1, G, dtype=np.uint64), len(groups)))
        parts.append((tok, sel, mut, seed * 1000 + t_idx, slot))
    rfp = np.full(n, -1, np.int32)
    base = (np.arange(n_groups, dtype=np.int32) * G)
    for j in range(1, G):
        rfp[base + j] = base
    return _assemble(parts, n, rfp, pinned, raw_keys)


def mixed_batch(state: SynthTrie, tokens, n_proofs: int, seed: int, p_mut: float = 0.02, pinned: bool = False):
    """config 5 shard: half account proofs, half storage proofs (each with its account in the same
    shard, i.e. pairs), p_mut of all proofs mutated."""
    return nested_batch(state, tokens, n_proofs // 2, seed, k_storage=1, p_excl=0.0, p_mut=p_mut, pinned=pinned)


def make_state_and_tokens(n_accounts: int, n_tokens: int, n_slots: int, seed: int):
    tokens = [SynthTrie(n_slots, seed * 7919 + 1 + i, kind=1) for i in range(n_tokens)]
    pool = np.concatenate([t.root for t in tokens]) if tokens else None
    state = SynthTrie(n_accounts, seed, kind=0, pool_roots=pool)
    return state, tokens


# ----------------------------------------------------------------------------- config 4: block tries
def block_tries(n_blocks: int, per_block: int = 300, kind: str = "tx", seed: int = 4, pinned: bool = False):
    """Synthetic per-block transaction (kind="tx") or receipt (kind="receipt") tries in the KvBatch
    layout, or both interleaved (kind="both": trie 2b = block b's transactions, 2b+1 = its receipts):
    key = rlp(index), value = opaque EIP-2718 bytes (first byte 0x02).  Sizes per SURVEY.md section 8d
    config 4: txs log-normal, median 180 B, capped at 8 KB; receipts median ~1.5 KB (256-byte bloom +
    0..40 logs), tail to 30 KB."""
    import zk_state_proofs_b200 as z
    rng = np.random.default_rng(seed + {"tx": 0, "receipt": 1000, "both": 2000}[kind])
    n_tries = n_blocks * (2 if kind == "both" else 1)
    n = n_tries * per_block
    tx_lens = lambda k: np.clip(rng.lognormal(np.log(180.0), 0.9, k), 100, 8192).astype(np.uint32)
    rc_lens = lambda k: np.clip(270 + rng.lognormal(np.log(1230.0), 1.0, k), 270, 30000).astype(np.uint32)
    if kind == "tx":
        lens = tx_lens(n)
    elif kind == "receipt":
        lens = rc_lens(n)
    else:
        lens = np.stack([tx_lens(n // 2).reshape(n_blocks, per_block), rc_lens(n // 2).reshape(n_blocks, per_block)],
                        axis=1).reshape(-1)
    padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
    value_off = np.zeros(n, np.uint64)
    np.cumsum(padded[:-1], out=value_off[1:])
    total = int(padded.sum()) + 16
    value_bytes = _alloc(total, np.uint8, pinned)
    step = 1 << 28
    for s in range(0, total, step):
        e = min(total, s + step)
        value_bytes[s:e] = rng.integers(0, 256, e - s, dtype=np.uint8)
    value_bytes[value_off.astype(np.int64)] = 2
    one = b"".join(z.rlp_index(i) for i in range(per_block))
    klen = np.array([len(z.rlp_index(i)) for i in range(per_block)], np.int64)
    key_bytes = np.frombuffer(one * n_tries + b"\0" * 16, np.uint8).copy()
    key_off = np.zeros(n + 1, np.uint32)
    np.cumsum(np.tile(klen, n_tries), out=key_off[1:])
    trie_first = (np.arange(n_tries + 1, dtype=np.uint64) * per_block).astype(np.uint32)
    return z.KvBatch(key_bytes, key_off, value_bytes, value_off, lens, trie_first)


def batch_to_borsh(b, pinned: bool = False):
    """A CSR batch as borsh(MerkleProofInput) blobs (one uint8 array + [n + 1] uint64 offsets), the format a
    prover's input file holds (crypto-ops/src/types.rs:4-9); input of mptv_verify_borsh / mptv_flatten_borsh."""
    L = lib()
    n = b.n_proofs
    off = np.zeros(n + 1, np.uint64)
    args = [b.node_bytes.ctypes.data, b.node_off.ctypes.data, b.node_len.ctypes.data, b.proof_first.ctypes.data, n,
            b.roots.ctypes.data, b.key_bytes.ctypes.data, b.key_off.ctypes.data, off.ctypes.data]
    total = int(L.mptgen_csr_to_borsh(*args, None, 1))
    blobs = _alloc(total + 16, np.uint8, pinned)
    L.mptgen_csr_to_borsh(*args, blobs.ctypes.data, _threads())
    return blobs[:total], off


def batch_to_storage_borsh(b, pinned: bool = False):
    """A nested CSR batch made with raw_keys=True as borsh(StorageProofInput) blobs (crypto-ops/src/types.rs:11-19), one per
    group -- the storage guest's input (storage-circuit/src/main.rs:6-9); input of mptv_verify_storage_borsh.
    -> (blobs u8[], blob_off u64[n_groups + 1], group_first u64[n_groups + 1])"""
    L = lib()
    n = b.n_proofs
    gf = np.zeros(n + 1, np.uint64)
    off = np.zeros(n + 1, np.uint64)
    ng = ctypes.c_uint64()
    rfp = np.ascontiguousarray(b.root_from_proof, np.int32)
    args = [b.node_bytes.ctypes.data, b.node_off.ctypes.data, b.node_len.ctypes.data, b.proof_first.ctypes.data, n,
            b.roots.ctypes.data, b.key_bytes.ctypes.data, b.raw_keys.ctypes.data, rfp.ctypes.data, gf.ctypes.data,
            ctypes.byref(ng), off.ctypes.data]
    total = int(L.mptgen_csr_to_storage_borsh(*args, None, 1))
    blobs = _alloc(total + 16, np.uint8, pinned)
    L.mptgen_csr_to_storage_borsh(*args, blobs.ctypes.data, _threads())
    g = int(ng.value)
    return blobs[:total], off[:g + 1].copy(), gf[:g + 1].copy()
