"""oracle/fuzzgen.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Deterministic (seeded) generators of verify_merkle_proof inputs (root, proof, key):
well-formed tries with inclusion / exclusion proofs, the config-3 mutators (bit flips,
dropped nodes, wrong key, shuffles, junk), and structure-aware malformed nodes that
exercise every rule of SURVEY.md Appendix A (non-canonical RLP, bad child lengths,
hex-prefix flags > 3, empty paths, list-typed values, 0x81 XX values, trailing bytes,
inline children of every kind, at the root and below it).

Each case is a dict(root=bytes, proof=[bytes], key=bytes, tag=str).  The expected
outcome is NOT computed here: it comes from the reference ELF (oracle/gen_golden.py,
tests/test_oracle_vs_reference_elf.py).
"""
from __future__ import annotations

import random

from .pytrie import Trie, hex_prefix, nibbles, rlp_hdr, rlp_list, rlp_str


# ------------------------------------------------------------------ valid tries
def random_trie(rng: random.Random, keccak, kind: str):
    """-> (Trie, kv dict)"""
    kv = {}
    if kind == "state":  # 32-byte hashed keys, account-sized values
        n = rng.choice([1, 2, 3, 5, 17, 40, 120])
        for _ in range(n):
            kv[rng.randbytes(32)] = rng.randbytes(rng.randint(60, 110))
    elif kind == "storage":  # 32-byte keys, tiny values (incl. single bytes both sides of 0x80)
        n = rng.choice([1, 2, 4, 20, 60])
        for _ in range(n):
            v = rng.choice([bytes([rng.randint(1, 0x7F)]), bytes([rng.randint(0x80, 0xFF)]),
                            rng.randbytes(rng.randint(2, 33))])
            kv[rng.randbytes(32)] = rlp_str(v.lstrip(b"\x00") or b"\x01")
    elif kind == "tx":  # keys rlp(index)
        n = rng.choice([1, 2, 16, 130, 200])
        big = rng.random() < 0.5
        from .pytrie import rlp_uint
        for i in range(n):
            ln = rng.randint(100, 300) if big else rng.randint(1, 12)
            kv[rlp_uint(i)] = b"\x02" + rng.randbytes(ln)
    elif kind == "short":  # short keys with prefix relations -> branch values, extensions, inline nodes
        n = rng.choice([2, 3, 6, 12])
        alphabet = [0x12, 0x13, 0x22, 0x34, 0x10, 0x00, 0xFF]
        for _ in range(n):
            k = bytes(rng.choice(alphabet) for _ in range(rng.randint(0, 4)))
            kv[k] = rng.randbytes(rng.choice([1, 1, 2, 5, 31, 32, 33, 40]))
    else:
        raise ValueError(kind)
    return Trie(kv, keccak), kv


def valid_cases(rng: random.Random, keccak, n_tries: int):
    out = []
    for _ in range(n_tries):
        kind = rng.choice(["state", "storage", "tx", "short"])
        t, kv = random_trie(rng, keccak, kind)
        keys = list(kv.keys())
        for k in rng.sample(keys, min(3, len(keys))):
            out.append(dict(root=t.root, proof=t.proof(k), key=k, tag=f"{kind}/incl"))
        # absent keys: random, and near-misses of a present key
        for _ in range(2):
            k0 = rng.choice(keys)
            choices = [rng.randbytes(len(k0) or 1), k0[:-1], k0 + b"\x00"]
            if k0:
                kb = bytearray(k0)
                kb[rng.randrange(len(kb))] ^= 1 << rng.randrange(8)
                choices.append(bytes(kb))
            k = rng.choice(choices)
            out.append(dict(root=t.root, proof=t.proof(k), key=k, tag=f"{kind}/absent"))
    return out


# ---------------------------------------------------------------------- mutators
def mutate(rng: random.Random, case: dict, keccak):
    """One of the config-3 mutators plus a few more."""
    proof = [bytes(p) for p in case["proof"]]
    root, key = case["root"], case["key"]
    m = rng.choice(["flip_leaf", "flip_inner", "flip_any", "drop_last", "drop_root", "drop_mid",
                    "wrong_key", "shuffle", "junk", "dup", "trunc_node", "extend_node",
                    "wrong_root", "short_key", "long_key", "swap_root", "empty"])
    if m in ("flip_leaf", "flip_inner", "flip_any") and proof:
        i = len(proof) - 1 if m == "flip_leaf" else (rng.randrange(max(1, len(proof) - 1)) if m == "flip_inner"
                                                      else rng.randrange(len(proof)))
        b = bytearray(proof[i])
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        proof[i] = bytes(b)
    elif m == "drop_last" and proof:
        proof.pop()
    elif m == "drop_root" and proof:
        proof.pop(0)
    elif m == "drop_mid" and len(proof) > 2:
        proof.pop(rng.randrange(1, len(proof) - 1))
    elif m == "wrong_key":
        key = rng.randbytes(len(key) or 1)
    elif m == "shuffle":
        rng.shuffle(proof)
    elif m == "junk":
        j = rng.choice([b"\xc0", b"", b"\x80", rng.randbytes(rng.randint(1, 80)),
                        rlp_list([rlp_str(rng.randbytes(3))] * 3)])
        proof.insert(rng.randrange(len(proof) + 1), j)
    elif m == "dup" and proof:
        proof.insert(rng.randrange(len(proof) + 1), rng.choice(proof))
    elif m == "trunc_node" and proof:
        i = rng.randrange(len(proof))
        proof[i] = proof[i][:rng.randrange(len(proof[i]) + 1)]
    elif m == "extend_node" and proof:
        i = rng.randrange(len(proof))
        proof[i] = proof[i] + rng.randbytes(rng.randint(1, 3))
    elif m == "wrong_root":
        r = bytearray(root)
        r[rng.randrange(32)] ^= 1 << rng.randrange(8)
        root = bytes(r)
    elif m == "short_key" and key:
        key = key[:-1]
    elif m == "long_key":
        key = key + bytes([rng.randrange(256)])
    elif m == "swap_root" and proof:
        # make a mutated root node the new root (keeps hash link valid): exercises root-only rules
        b = bytearray(proof[0])
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        proof[0] = bytes(b)
        root = keccak(proof[0])
    elif m == "empty":
        proof = []
    return dict(root=root, proof=proof, key=key, tag=case["tag"] + "+" + m)


# -------------------------------------------------- structure-aware malformed nodes
def _weird_header(rng, n, is_list):
    """Possibly non-canonical header for a payload of n bytes."""
    base = 0xC0 if is_list else 0x80
    r = rng.random()
    if r < 0.75:
        return rlp_hdr(n, is_list)
    if r < 0.85:  # long form for a short payload (NonCanonicalSize)
        return bytes([base + 56, n & 0xFF])
    if r < 0.92:  # leading zero in the length
        return bytes([base + 57, 0, n & 0xFF])
    if r < 0.96:  # claims more than there is
        return rlp_hdr(n + rng.randint(1, 40), is_list)
    return rlp_hdr(max(0, n - 1), is_list)  # claims less


def _inline_leaf(rng, good=True):
    flag = rng.choice([2, 3]) if good else rng.choice([4, 5, 8, 15, 0, 1])
    nib = [rng.randrange(16) for _ in range(rng.randint(0, 3))]
    hp = bytearray(hex_prefix(nib, True))
    hp[0] = (hp[0] & 0x0F) | ((flag | (len(nib) & 1)) << 4) if good else (hp[0] & 0x0F) | (flag << 4)
    val = rng.choice([b"\x05", b"\xcc", b"", rng.randbytes(2), rng.randbytes(8), rng.randbytes(40)])
    return rlp_list([rlp_str(bytes(hp)), rlp_str(val)])


def weird_item(rng, keccak):
    r = rng.randrange(22)
    if r == 0: return b"\x80"
    if r == 1: return rlp_str(rng.randbytes(32))
    if r == 2: return rlp_str(rng.randbytes(31))
    if r == 3: return rlp_str(rng.randbytes(33))
    if r == 4: return bytes([rng.randrange(0x80)])
    if r == 5: return bytes([0x81, rng.randrange(0x80, 0x100)])
    if r == 6: return bytes([0x81, rng.randrange(0x80)])  # non-canonical single byte
    if r == 7: return _inline_leaf(rng, True)
    if r == 8: return _inline_leaf(rng, False)
    if r == 9: return b"\xc0"
    if r == 10: return rlp_list([rlp_str(rng.randbytes(2))] * 3)
    if r == 11: return rlp_str(rng.randbytes(rng.randint(56, 70)))
    if r == 12:  # inline extension -> hash
        return rlp_list([rlp_str(hex_prefix([rng.randrange(16)], False)), rlp_str(rng.randbytes(32))])
    if r == 13:  # inline extension with empty / bad child
        return rlp_list([rlp_str(hex_prefix([1, 2], False)), rng.choice([b"\x80", rlp_str(rng.randbytes(5))])])
    if r == 14:  # empty hex-prefix path item
        return rlp_list([b"\x80", rlp_str(rng.randbytes(3))])
    if r == 15:  # inline branch (small)
        return rlp_list([b"\x80"] * 16 + [rlp_str(rng.randbytes(rng.randint(0, 4)))])
    if r == 16:  # list-typed value inside an inline leaf
        return rlp_list([rlp_str(hex_prefix([3], True)), rlp_list([rlp_str(rng.randbytes(4))])])
    if r == 17:  # weird header string
        pl = rng.randbytes(rng.randint(0, 40))
        return _weird_header(rng, len(pl), False) + pl
    if r == 18:  # weird header list
        pl = _inline_leaf(rng, True)
        return _weird_header(rng, len(pl), True) + pl
    if r == 19: return rlp_str(rng.randbytes(rng.randint(2, 30)))
    if r == 20:  # list as path item
        return rlp_list([rlp_list([b"\x20"]), rlp_str(rng.randbytes(3))])
    return rng.randbytes(rng.randint(1, 6))


def weird_node(rng, keccak, force_len=None):
    """A list node with a chosen number of items, most of them plausible."""
    k = force_len if force_len is not None else rng.choice([17] * 8 + [2] * 6 + [0, 1, 3, 16, 18])
    items = []
    if k in (17, 16, 18):
        for _ in range(k - (1 if k >= 17 else 0)):
            items.append(weird_item(rng, keccak) if rng.random() < 0.15 else
                         rng.choice([b"\x80", b"\x80", rlp_str(rng.randbytes(32)), _inline_leaf(rng, True)]))
        if k >= 17:
            items.append(rng.choice([b"\x80", b"\x80", rlp_str(rng.randbytes(rng.randint(1, 40))),
                                     bytes([0x81, 0xEE]), b"\x07", rlp_list([b"\x01", b"\x02"]),
                                     weird_item(rng, keccak)]))
    elif k == 2:
        leaf = rng.random() < 0.6
        nib = [rng.randrange(16) for _ in range(rng.randint(0, 5))]
        path = rlp_str(hex_prefix(nib, leaf))
        if rng.random() < 0.3:
            path = weird_item(rng, keccak)
        if rng.random() < 0.15:  # non-zero pad nibble on an even path
            hp = bytearray(hex_prefix(nib, leaf))
            if not len(nib) & 1:
                hp[0] |= rng.randrange(1, 16)
            path = rlp_str(bytes(hp))
        if leaf:
            second = rng.choice([rlp_str(rng.randbytes(rng.randint(0, 50))), bytes([0x81, 0xF0]), b"\x33",
                                 rlp_list([rlp_str(rng.randbytes(10))]), weird_item(rng, keccak)])
        else:
            second = rng.choice([rlp_str(rng.randbytes(32)), _inline_leaf(rng, True), b"\x80",
                                 weird_item(rng, keccak)])
        items = [path, second]
    else:
        items = [weird_item(rng, keccak) for _ in range(k)]
    pl = b"".join(items)
    node = (_weird_header(rng, len(pl), True) if rng.random() < 0.15 else rlp_hdr(len(pl), True)) + pl
    if rng.random() < 0.1:
        node += rng.randbytes(rng.randint(1, 4))
    if rng.random() < 0.03:
        node = rng.choice([b"\x80", rlp_str(rng.randbytes(32)), rlp_str(rng.randbytes(7)), b"", b"\x05"])
    return node, items


def _key_into(rng, node_items, prefix_nibbles):
    """A key (bytes) whose nibble path starts with prefix_nibbles and then wanders."""
    nib = list(prefix_nibbles) + [rng.randrange(16) for _ in range(rng.randint(0, 6))]
    if len(nib) & 1:
        nib.append(rng.randrange(16))
    return bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))


def weird_cases(rng: random.Random, keccak, n: int):
    out = []
    for _ in range(n):
        mode = rng.random()
        if mode < 0.45:  # weird node IS the root
            node, items = weird_node(rng, keccak)
            root = keccak(node)
            proof = [node]
            # also supply hash-referenced children sometimes
            extra = _inline_leaf(rng, True) + rng.randbytes(0)
            key = _key_into(rng, items, [rng.randrange(16)])
            if rng.random() < 0.3:
                key = b""
            out.append(dict(root=root, proof=proof, key=key, tag="weird/root"))
        else:  # valid root branch -> hash -> weird child (>= 32 bytes most of the time)
            child, items = weird_node(rng, keccak)
            slot = rng.randrange(16)
            slots = [b"\x80"] * 16
            slots[slot] = rlp_str(keccak(child))
            other = rng.randrange(16)
            if other != slot and rng.random() < 0.5:
                slots[other] = rlp_str(rng.randbytes(32))  # dangling hash in an untouched slot
            rootn = rlp_list(slots + [b"\x80"])
            # third level sometimes: if the child is a branch/ext with a hash we know, provide a leaf for it
            proof = [rootn, child]
            grand = rlp_list([rlp_str(hex_prefix([rng.randrange(16) for _ in range(3)], True)),
                              rlp_str(rng.randbytes(40))])
            if rng.random() < 0.3:
                proof.append(grand)
            key = _key_into(rng, items, [slot, rng.randrange(16)])
            out.append(dict(root=keccak(rootn), proof=proof, key=key, tag="weird/child"))
    return out


def _nested(rng, depth, used, corrupt):
    """An inline node nested `depth` levels deep.  -> (encoding, nibble path to its value, value).
    `used` = nibbles consumed above (the leaf picks its path length so that the whole key is byte aligned);
    `corrupt` = probability of planting one decode fault (bad hex-prefix flag, 3-item list, 31-byte child
    string, empty extension path) somewhere on or beside the path."""
    kind = rng.choice(["ext", "branch", "branch", "leaf"]) if depth > 0 else "leaf"
    if kind == "leaf":
        n = rng.randint(0, 4)
        if (used + n) & 1:
            n += 1
        nib = [rng.randrange(16) for _ in range(n)]
        val = rng.choice([b"\x05", b"\xcc", b"", rng.randbytes(3), rng.randbytes(20), rng.randbytes(60)])
        hp = bytearray(hex_prefix(nib, True))
        if rng.random() < corrupt:
            hp[0] = (hp[0] & 0x0F) | (rng.choice([4, 7, 12]) << 4)
        items = [rlp_str(bytes(hp)), rlp_str(val)]
        if rng.random() < corrupt:
            items.append(b"\x01")
        return rlp_list(items), nib, val
    if kind == "ext":
        nib = [rng.randrange(16) for _ in range(rng.randint(1, 3))]
        if rng.random() < corrupt * 0.5:
            nib = []
        child, path, val = _nested(rng, depth - 1, used + len(nib), corrupt)
        return rlp_list([rlp_str(hex_prefix(nib, False)), child]), nib + path, val
    slot = rng.randrange(16)
    child, path, val = _nested(rng, depth - 1, used + 1, corrupt)
    slots = []
    for i in range(16):
        if i == slot:
            slots.append(child)
        else:
            r = rng.random()
            if r < 0.80: slots.append(b"\x80")
            elif r < 0.88: slots.append(rlp_str(rng.randbytes(32)))
            elif r < 0.96: slots.append(_inline_leaf(rng, True))
            elif r < 0.96 + corrupt: slots.append(rlp_str(rng.randbytes(rng.choice([31, 33, 5]))))
            else: slots.append(_nested(rng, min(depth - 1, 2), 0, corrupt)[0])
    bval = b"\x80" if rng.random() < 0.9 else rlp_str(rng.randbytes(rng.randint(1, 9)))
    return rlp_list(slots + [bval]), [slot] + path, val


def nested_cases(rng: random.Random, keccak, n: int):
    """Inline nodes inside inline nodes, up to 12 levels, as the root node and below a hashed root, with the
    key that leads to the innermost value (or a neighbour of it) and an occasional planted decode fault."""
    out = []
    for _ in range(n):
        depth = rng.randint(1, 12)
        corrupt = rng.choice([0.0, 0.0, 0.03, 0.1])
        as_root = rng.random() < 0.4
        node, path, _ = _nested(rng, depth, 0 if as_root else 1, corrupt)
        if as_root:
            proof, root, nib = [node], keccak(node), path
        else:
            slot = rng.randrange(16)
            slots = [b"\x80"] * 16
            slots[slot] = rlp_str(keccak(node)) if len(node) >= 32 or rng.random() < 0.5 else node
            rootn = rlp_list(slots + [b"\x80"])
            proof, root, nib = [rootn, node], keccak(rootn), [slot] + path
        if len(nib) & 1:
            nib = nib + [rng.randrange(16)]
        key = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
        r = rng.random()
        if r < 0.15 and key:       # a neighbouring key: leaves the path somewhere inside the nesting
            k = bytearray(key); i = rng.randrange(len(k)); k[i] ^= 1 << rng.randrange(8); key = bytes(k)
        elif r < 0.22:
            key = key[:-1]
        elif r < 0.28:
            key = key + bytes([rng.randrange(256)])
        if rng.random() < 0.2:
            rng.shuffle(proof)
        out.append(dict(root=root, proof=proof, key=key, tag="nested/root" if as_root else "nested/child"))
    return out


def deep_nested_cases(rng: random.Random, keccak, n: int, depths=(64, 65, 100, 128, 129, 200, 500, 1000, 2000)):
    """Inline nodes nested 64 ... 2 000 levels deep (built bottom-up, no recursion): mostly inline branches -- the
    levels that need a decode frame each -- with inline extensions mixed in (the decoder's tail calls), decorated
    siblings (hash references, inline leaves, small nested branches) and, with probability `corrupt`, ONE decode
    fault planted at a random level in a slot before or after the nested child, so that it is met on the way down
    or only after the decoder has unwound from the full depth below it.  The reference accepts any depth."""
    out = []
    for it in range(n):
        depth = depths[it % len(depths)] + rng.choice([0, 0, 1, -1, 7])
        corrupt_level = rng.randrange(depth) if rng.random() < 0.45 else -1
        corrupt_after = rng.random() < 0.6
        as_root = rng.random() < 0.15
        # bottom leaf
        nlevels = depth
        kinds = [("ext" if rng.random() < 0.12 else "branch") for _ in range(nlevels)]
        ext_nibs = {lvl: [rng.randrange(16) for _ in range(rng.randint(1, 2))] for lvl in range(nlevels) if kinds[lvl] == "ext"}
        above = (0 if as_root else 1) + sum(1 for k in kinds if k == "branch") + sum(len(v) for v in ext_nibs.values())
        path_below = []  # nibbles from the current node down to the value
        n_leaf = rng.randint(0, 3)
        if (above + n_leaf) & 1:
            n_leaf += 1  # the whole key is byte aligned
        nib = [rng.randrange(16) for _ in range(n_leaf)]
        val = rng.choice([b"\x05", b"\xcc", b"", rng.randbytes(3), rng.randbytes(20)])
        hp = bytearray(hex_prefix(nib, True))
        if corrupt_level == -2:
            hp[0] = (hp[0] & 0x0F) | 0x40
        node = rlp_list([rlp_str(bytes(hp)), rlp_str(val)])
        path_below = list(nib)
        for lvl in range(nlevels - 1, -1, -1):
            if kinds[lvl] == "ext":
                en = ext_nibs[lvl]
                if lvl == corrupt_level and rng.random() < 0.3:
                    en = []  # empty extension path: a raw panic in the reference
                node = rlp_list([rlp_str(hex_prefix(en, False)), node])
                path_below = en + path_below
                continue
            slot = rng.randrange(16)
            slots = []
            for i in range(16):
                if i == slot:
                    slots.append(node)
                    continue
                r = rng.random()
                if lvl == corrupt_level and ((i > slot) == corrupt_after) and rng.random() < 0.5:
                    slots.append(rng.choice([rlp_str(rng.randbytes(31)), rlp_str(rng.randbytes(33)), rlp_list([b"\x01", b"\x02", b"\x03"]),
                                             rlp_list([rlp_str(b"\x45\x01"), rlp_str(b"v")]), rlp_list([rlp_str(b""), rlp_str(b"v")])]))
                elif r < 0.93: slots.append(b"\x80")
                elif r < 0.96: slots.append(rlp_str(rng.randbytes(32)))
                elif r < 0.99: slots.append(_inline_leaf(rng, True))
                else: slots.append(_nested(rng, 2, 0, 0.0)[0])
            bval = b"\x80" if rng.random() < 0.97 else rlp_str(rng.randbytes(rng.randint(1, 9)))
            node = rlp_list(slots + [bval])
            path_below = [slot] + path_below
        if as_root:
            proof, root, nibs = [node], keccak(node), path_below
        else:
            slot = rng.randrange(16)
            slots = [b"\x80"] * 16
            slots[slot] = rlp_str(keccak(node))
            rootn = rlp_list(slots + [b"\x80"])
            proof, root, nibs = [rootn, node], keccak(rootn), [slot] + path_below
        if len(nibs) & 1:
            nibs = nibs + [rng.randrange(16)]
        key = bytes((nibs[i] << 4) | nibs[i + 1] for i in range(0, len(nibs), 2))
        r = rng.random()
        if r < 0.12 and key:   # a neighbouring key: leaves the path somewhere inside the nesting
            k = bytearray(key); i = rng.randrange(len(k)); k[i] ^= 1 << rng.randrange(8); key = bytes(k)
        elif r < 0.16:
            key = key[:-1]
        out.append(dict(root=root, proof=proof, key=key, tag=f"deep/{'root' if as_root else 'child'}-{depth}"))
    return out


def resealed_cases(rng: random.Random, keccak, base, n: int):
    """Corrupt ONE node of a valid root-first proof at the byte level (bit flips, inserted / deleted / overwritten
    bytes, truncation) and then RE-SEAL the chain: every ancestor gets the corrupted child's new hash and the
    root hash is recomputed, so all hash links hold and the reference really decodes and walks the corrupted
    node (a plain flip only ever yields a dangling link)."""
    out = []
    chains = [c for c in base if len(c["proof"]) >= 1 and c["tag"].endswith(("/incl", "/absent"))]
    while len(out) < n and chains:
        c = rng.choice(chains)
        proof = [bytes(p) for p in c["proof"]]
        i = rng.randrange(len(proof))
        b = bytearray(proof[i])
        for _ in range(rng.choice([1, 1, 1, 2, 3])):
            op = rng.choice(["flip", "flip", "set", "ins", "del", "trunc", "hdr"])
            if not b:
                break
            j = rng.randrange(len(b))
            if op == "flip": b[j] ^= 1 << rng.randrange(8)
            elif op == "set": b[j] = rng.choice([0x00, 0x7f, 0x80, 0x81, 0xa0, 0xb7, 0xb8, 0xc0, 0xc1, 0xf7, 0xf8, 0xff])
            elif op == "ins": b.insert(j, rng.randrange(256))
            elif op == "del": del b[j]
            elif op == "trunc": del b[rng.randrange(len(b)):]
            else: b[rng.randrange(min(4, len(b)))] = rng.randrange(256)   # aim at the list / first item header
        old = proof[i]
        proof[i] = bytes(b)
        ok = True
        for j in range(i - 1, -1, -1):   # propagate the new hash up the chain
            oh, nh = keccak(old), keccak(proof[j + 1])
            if len(old) < 32 or oh not in proof[j]:
                ok = False   # the child was inline / not referenced by hash: skip this one
                break
            old = proof[j]
            proof[j] = proof[j].replace(oh, nh, 1)
        if not ok:
            continue
        root = keccak(proof[0])   # the re-sealed root (the chain is root first)
        key = c["key"]
        if rng.random() < 0.1:
            key = key[:-1] if rng.random() < 0.5 else key + bytes([rng.randrange(256)])
        if rng.random() < 0.1:
            rng.shuffle(proof)
        out.append(dict(root=root, proof=proof, key=key, tag=c["tag"] + f"+reseal{i}"))
    return out


def padded_cases(rng: random.Random, cases, n: int):
    """Long proofs: a case's own nodes scattered among up to 100 junk strings, nodes of OTHER cases and
    duplicates of its own nodes (the reference accepts any multiset of nodes, lib.rs:10-13)."""
    pool = [nd for c in cases for nd in c["proof"] if nd]
    out = []
    for _ in range(n):
        c = rng.choice(cases)
        proof = list(c["proof"])
        for _ in range(rng.choice([3, 7, 9, 15, 17, 31, 33, 60, 100])):
            r = rng.random()
            extra = rng.randbytes(rng.randint(1, 80)) if r < 0.4 else (rng.choice(pool) if r < 0.8 or not proof
                                                                       else rng.choice(proof))
            proof.insert(rng.randrange(len(proof) + 1), extra)
        if rng.random() < 0.5:
            rng.shuffle(proof)
        out.append(dict(root=c["root"], proof=proof, key=c["key"], tag="padded/" + c["tag"]))
    return out


def corpus(seed: int, keccak, n_tries: int, n_mut: int, n_weird: int, n_nested: int = 0, n_reseal: int = 0,
           n_padded: int = 0):
    rng = random.Random(seed)
    base = valid_cases(rng, keccak, n_tries)
    out = list(base)
    for _ in range(n_mut):
        out.append(mutate(rng, rng.choice(base), keccak))
    out += weird_cases(rng, keccak, n_weird)
    if n_nested:
        out += nested_cases(rng, keccak, n_nested)
    if n_reseal:
        out += resealed_cases(rng, keccak, base, n_reseal)
    if n_padded:  # drawn last, so corpora made without it are unchanged
        out += padded_cases(rng, list(out), n_padded)
    return out
