"""oracle/pytrie.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A tiny pure-Python Merkle-Patricia-Trie builder / prover (Yellow-Paper appendix D
definition), used by tests and oracle/gen_golden.py to make SMALL tries whose proofs
are then judged by the reference ELF, the C restatement and the CUDA path.  It stands
in for what trie-utils does with eth_trie (EthTrie::new / insert / root_hash /
get_proof, /root/reference/trie-utils/src/proofs/transaction.rs:41-68) on a handful
of keys.  Pure-Python loops: small cases only.
"""
from __future__ import annotations


def rlp_str(b: bytes) -> bytes:
    if len(b) == 1 and b[0] < 0x80:
        return bytes(b)
    return rlp_hdr(len(b), False) + bytes(b)


def rlp_hdr(n: int, is_list: bool) -> bytes:
    base = 0xC0 if is_list else 0x80
    if n < 56:
        return bytes([base + n])
    be = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([base + 55 + len(be)]) + be


def rlp_list(items) -> bytes:
    pl = b"".join(items)
    return rlp_hdr(len(pl), True) + pl


def rlp_uint(v: int) -> bytes:
    if v == 0:
        return b"\x80"
    return rlp_str(v.to_bytes((v.bit_length() + 7) // 8, "big"))


def nibbles(key: bytes):
    out = []
    for b in key:
        out += [b >> 4, b & 15]
    return out


def hex_prefix(nib, leaf: bool) -> bytes:
    flag = 2 if leaf else 0
    if len(nib) & 1:
        out = [((flag | 1) << 4) | nib[0]]
        rest = nib[1:]
    else:
        out = [flag << 4]
        rest = nib
    for i in range(0, len(rest), 2):
        out.append((rest[i] << 4) | rest[i + 1])
    return bytes(out)


class Trie:
    """Build from a dict {key bytes: value bytes}; empty values are skipped (eth_trie's
    insert(key, b"") is a delete)."""

    def __init__(self, kv: dict, keccak):
        self.keccak = keccak
        items = sorted((tuple(nibbles(k)), v) for k, v in kv.items() if len(v) > 0)
        self.nodes = {}  # hash -> encoded
        self.root_enc = self._build(items, 0) if items else b"\x80"
        self.root = keccak(self.root_enc)

    # returns the encoded node for items (all sharing the first `depth` nibbles)
    def _build(self, items, depth) -> bytes:
        if len(items) == 1:
            k, v = items[0]
            return rlp_list([rlp_str(hex_prefix(list(k[depth:]), True)), rlp_str(v)])
        # common prefix beyond depth
        first, last = items[0][0], items[-1][0]
        cp = 0
        while depth + cp < len(first) and depth + cp < len(last) and first[depth + cp] == last[depth + cp]:
            cp += 1
        if cp > 0:
            child = self._build(items, depth + cp)
            return rlp_list([rlp_str(hex_prefix(list(first[depth:depth + cp]), False)), self._ref(child)])
        slots = []
        value = b""
        rest = items
        if len(items[0][0]) == depth:
            value = items[0][1]
            rest = items[1:]
        for nib in range(16):
            grp = [it for it in rest if it[0][depth] == nib]
            slots.append(self._ref(self._build(grp, depth + 1)) if grp else b"\x80")
        slots.append(rlp_str(value))
        return rlp_list(slots)

    def _ref(self, enc: bytes) -> bytes:
        if len(enc) < 32:
            return enc
        h = self.keccak(enc)
        self.nodes[h] = enc
        return rlp_str(h)

    def proof(self, key: bytes):
        """Nodes on the path of `key`, root first: the root always, other nodes only when
        referenced by hash (>= 32 bytes), like eth_trie's get_proof."""
        from .pyrlp import split_list  # local import to keep module load light
        if self.root_enc == b"\x80":
            return []  # empty trie: eth_trie's get_proof pushes the root only when it is not Node::Empty
        out = [self.root_enc]
        path = nibbles(key) + [16]
        enc = self.root_enc
        idx = 0
        while True:
            items = split_list(enc)
            if items is None:
                break
            if len(items) == 17:
                if idx >= len(path) or path[idx] == 16:
                    break
                child = items[path[idx]]
                idx += 1
            elif len(items) == 2:
                hp = payload(items[0])
                flag = hp[0] >> 4
                nib = ([hp[0] & 15] if flag & 1 else []) + nibbles(hp[1:])
                if flag >= 2:
                    break
                if path[idx:idx + len(nib)] != nib:
                    break
                idx += len(nib)
                child = items[1]
            else:
                break
            if child == b"\x80":
                break
            if child[0] >= 0xC0:
                enc = child  # inline node: part of the parent, not a separate proof node
                continue
            h = payload(child)
            enc = self.nodes[h]
            out.append(enc)
        return out


def payload(item: bytes) -> bytes:
    b = item[0]
    if b < 0x80:
        return item[:1]
    if b < 0xB8:
        return item[1:1 + b - 0x80]
    if b < 0xC0:
        ll = b - 0xB7
        return item[1 + ll:]
    if b < 0xF8:
        return item[1:1 + b - 0xC0]
    ll = b - 0xF7
    return item[1 + ll:]
