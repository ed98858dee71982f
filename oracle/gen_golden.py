"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/verify_vectors.json.gz by
running the REFERENCE ITSELF (its committed SP1 guest ELF under oracle/rv32emu.c) on a seeded
corpus (oracle/fuzzgen.py) plus the 14 known-answer vectors of SURVEY.md Appendix E.

Only runnable where /root/reference is mounted (this container).  The committed output lets the
GPU box -- which has no /root/reference -- check the C restatement and the CUDA path against
results produced by the reference's own code.

    python -m oracle.gen_golden
"""
import gzip
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

from .fuzzgen import corpus
from .pyoracle import Oracle, RefElf

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "verify_vectors.json.gz")

APPENDIX_E = [
    (["cc822080880102030405060708"], "0e9985286c0f4a35519eeb86fa50ce8134ed1fd1bb88e6f74a9e4cb6f505079c", "80"),
    (["cc822080880102030405060708"], "0e9985286c0f4a35519eeb86fa50ce8134ed1fd1bb88e6f74a9e4cb6f505079c", "01"),
    (["c582208081ff"], "597b7d6dac7ed0e172717eb3ecee0ebd56188817afe42be254e47a5d322efc89", "80"),
    (["c482208080"], "1e03594df303045ca22e7d8f7ff4504ea2a9e1168864bf7b16ec37b79d9e2671", "80"),
    (["cc822f80880102030405060708"], "95f9dad6c16207e5c5c01f9b4b3867ed8db2e1e0f49e864633875b422902607b", "80"),
    (["d780c43082aabbc230058080808080808080808080808080"], "3888c5866c792987e82c5b40231486a304f99b1fce762a1f5a6fd9e798f990ba", "10"),
    (["d780c43082aabbc230058080808080808080808080808080"], "3888c5866c792987e82c5b40231486a304f99b1fce762a1f5a6fd9e798f990ba", "20"),
    (["d780c43082aabbc230058080808080808080808080808080"], "3888c5866c792987e82c5b40231486a304f99b1fce762a1f5a6fd9e798f990ba", "30"),
    (["d780c43082aabbc230058080808080808080808080808080"], "3888c5866c792987e82c5b40231486a304f99b1fce762a1f5a6fd9e798f990ba", "1000"),
    (["d880c43082aabbc33081cc8080808080808080808080808080"], "5feb4ffa44f09cd536c031d463bec819bae2ae0bbb16c111d391d5b456cfdbf1", "10"),
    (["e43ca2404142434445464748494a4b4c4d4e4f505152535455565758595a5b5c5d5e5f6061", "c0",
      "f180808080808080808080a03f68b9461a405e7e99ba7734a1836f01ac1f95840617f8f8d6bcde48dbb0f9ef808080808080"],
     "48245dc08200ffbefb3f22a59d0b178b5f4bb15fe07545be92e30723098841b5", "ac"),
    (["f180808080808080808080a03f68b9461a405e7e99ba7734a1836f01ac1f95840617f8f8d6bcde48dbb0f9ef808080808080"],
     "48245dc08200ffbefb3f22a59d0b178b5f4bb15fe07545be92e30723098841b5", "ac"),
    (["f180808080808080808080a03f68b9461a405e7e99ba7734a1836f01ac1f95840617f8f8d6bcde48dbb0f9ef808080808080",
      "e43ca2404142434445464748494a4b4c4d4e4f505152535455565758595a5b5c5d5e5f6060"],
     "48245dc08200ffbefb3f22a59d0b178b5f4bb15fe07545be92e30723098841b5", "ac"),
    (["f180808080808080808080a03f68b9461a405e7e99ba7734a1836f01ac1f95840617f8f8d6bcde48dbb0f9ef808080808080",
      "e43ca2404142434445464748494a4b4c4d4e4f505152535455565758595a5b5c5d5e5f6061"],
     "49245dc08200ffbefb3f22a59d0b178b5f4bb15fe07545be92e30723098841b5", "ac"),
]


def nested_cases(keccak):
    """Inline nodes nested far deeper than any real trie produces (an inline node is < 32 bytes): chains of
    inline extensions (2 bytes a level) and inline branches inside inline branches, as the root node (the
    lib.rs:19 re-encode assert fires: an inline child >= 32 bytes is not canonical) and below a hashed root
    (accepted by the reference at every depth: 64 ... 2 000 levels here; its guest ELF still answers at 8 000 and
    runs into its zkVM heap limit -- not a rule -- before 12 000)."""
    from .pytrie import hex_prefix, rlp_list, rlp_str
    nibs = [(i * 7 + 3) % 16 for i in range(128)]
    key = bytes((nibs[2 * i] << 4) | nibs[2 * i + 1] for i in range(64))
    out = []

    def under_root(node, tag):
        slots = [b"\x80"] * 16
        slots[nibs[0]] = rlp_str(keccak(node))
        root_node = rlp_list(slots + [b"\x80"])
        out.append(dict(root=keccak(root_node), proof=[root_node, node], key=key, tag=tag))

    for D in (1, 15, 16, 17, 40, 63, 64, 65, 100, 120):
        node = rlp_list([rlp_str(hex_prefix(nibs[1 + D:], True)), rlp_str(b"value-bytes-0123456789")])
        for k in range(D - 1, -1, -1):
            node = rlp_list([rlp_str(hex_prefix([nibs[1 + k]], False)), node])
        under_root(node, f"nested/ext-chain-{D}")
        if D in (1, 16, 64, 100):
            chain0 = rlp_list([rlp_str(hex_prefix(nibs[D:], True)), rlp_str(b"value-bytes-0123456789")])
            for k in range(D - 1, -1, -1):
                chain0 = rlp_list([rlp_str(hex_prefix([nibs[k]], False)), chain0])
            out.append(dict(root=keccak(chain0), proof=[chain0], key=key, tag=f"nested/ext-chain-as-root-{D}"))
    nibs64, key64 = nibs, key
    for B in (1, 2, 10, 40, 62, 63, 64, 65, 70, 127, 128, 129, 200, 500, 1000, 2000):
        if B + 2 > len(nibs64):   # a longer key for the nests that are deeper than 64 bytes of path
            nibs = [(i * 7 + 3) % 16 for i in range(2 * ((B + 6) // 2))]
            key = bytes((nibs[2 * i] << 4) | nibs[2 * i + 1] for i in range(len(nibs) // 2))
        else:
            nibs, key = nibs64, key64
        node = rlp_list([rlp_str(hex_prefix(nibs[1 + B:], True)), rlp_str(b"v")])
        for k in range(B - 1, -1, -1):
            slots = [b"\x80"] * 16
            slots[nibs[1 + k]] = node
            node = rlp_list(slots + [b"\x80"])
        # the top branch is frame 0 and the leaf sits B levels below it: beyond 63 the CUDA decoder's frame window
        # (kInlineWindow) wraps and is rebuilt by replay on the way back up
        under_root(node, f"nested/branch-in-branch-{B}")
    return out


def main():
    seed, n_tries, n_mut, n_weird = 7, 40, 500, 1200
    o = Oracle()
    ref = RefElf()
    cases = [dict(root=bytes.fromhex(r), proof=[bytes.fromhex(n) for n in ns], key=bytes.fromhex(k),
                  tag=f"appendixE/{i + 1}") for i, (ns, r, k) in enumerate(APPENDIX_E)]
    cases += corpus(seed, o.keccak256, n_tries, n_mut, n_weird)
    # a config-1 style case: 200-tx block trie, prove index 15 (tests/transaction.rs:13)
    import random
    from .pytrie import Trie, rlp_uint
    rng = random.Random(1)
    kv = {rlp_uint(i): b"\x02" + rng.randbytes(rng.randint(99, 299)) for i in range(200)}
    t = Trie(kv, o.keccak256)
    for i in (15, 0, 127, 128, 199, 200):
        cases.append(dict(root=t.root, proof=t.proof(rlp_uint(i)), key=rlp_uint(i), tag=f"config1/tx{i}"))

    cases += nested_cases(o.keccak256)
    import random as _random
    from .fuzzgen import nested_cases as fuzz_nested
    cases += [dict(c, tag="fuzz-" + c["tag"]) for c in fuzz_nested(_random.Random(99), o.keccak256, 600)]
    # byte-level corruption of one node with the hash chain re-sealed above it (the family that found the
    # "bare 32-byte string with trailing bytes panics" rule)
    from .fuzzgen import resealed_cases, valid_cases
    rr = _random.Random(123)
    cases += resealed_cases(rr, o.keccak256, valid_cases(rr, o.keccak256, 30), 300)
    # inline nodes nested 64 ... 2 000 levels deep with decorated siblings and planted decode faults
    from .fuzzgen import deep_nested_cases
    cases += [dict(c, tag="fuzz-" + c["tag"]) for c in deep_nested_cases(_random.Random(2024), o.keccak256, 180)]

    def one(c):
        return ref.run(c["root"], c["proof"], c["key"])

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, cases))
    vecs = []
    for c, r in zip(cases, res):
        vecs.append(dict(tag=c["tag"], root=c["root"].hex(), key=c["key"].hex(),
                         proof=[n.hex() for n in c["proof"]], status=r["status"],
                         value=None if r["value"] is None else r["value"].hex(), steps=r["steps"]))
    doc = dict(
        generator="oracle/gen_golden.py",
        reference_elf="circuits/elf/riscv32im-succinct-zkvm-elf (sha256 below)",
        corpus=dict(seed=seed, n_tries=n_tries, n_mut=n_mut, n_weird=n_weird),
        keccak_kats={"": o.keccak256(b"").hex(), "80": o.keccak256(b"\x80").hex()},
        vectors=vecs,
    )
    import hashlib
    doc["reference_elf_sha256"] = hashlib.sha256(ref.elf).hexdigest()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    from collections import Counter
    print(len(vecs), "vectors ->", OUT, os.path.getsize(OUT), "bytes")
    print(Counter(v["status"] for v in vecs))


if __name__ == "__main__":
    sys.exit(main())
