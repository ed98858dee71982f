"""oracle/gen_rebuild_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/rebuild_vectors.json.gz:
seeded key/value sets (tx-style rlp(i) keys, receipts with long values, inline leaves, duplicate keys,
deletes, keys that are prefixes of each other), their MPT roots, and for a few keys per trie the proof
of eth_trie's get_proof -- each proof JUDGED BY THE REFERENCE ITSELF (its guest ELF under
oracle/rv32emu.c): the reference accepts the proof against the root and returns the inserted bytes (or
reports the key absent), and its lib.rs:19 re-encode assert confirms the root node's canonical
encoding.  Roots are only recorded when the two independent builders (oracle/trie_oracle.c, sequential
insertion; oracle/pytrie.py, sorted recursion) agree.

Only runnable where /root/reference is mounted.      python -m oracle.gen_rebuild_golden
"""
import gzip
import json
import os
import random
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Oracle, RefElf  # noqa: E402
from oracle.pytrie import Trie  # noqa: E402
from tests.test_rebuild_oracle import make_kv, random_tries  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "rebuild_vectors.json.gz")


def main():
    o, ref = Oracle(), RefElf()
    tries = random_tries(2024, 64, sizes=(0, 1, 2, 3, 5, 17, 60, 130, 300))
    kv = make_kv(tries)
    roots, _, _ = o.trie_roots(kv, nthreads=4)
    rng = random.Random(7)
    jobs = []
    out_tries = []
    for t, kvs in enumerate(tries):
        final = dict(kvs)
        T = Trie(final, o.keccak256)
        assert T.root == roots[t].tobytes(), t
        keys = list(final)
        targets = rng.sample(keys, min(3, len(keys)))
        targets.append(bytes([rng.randrange(256) for _ in range(rng.choice([1, 2, 32]))]))
        ent = dict(items=[[k.hex(), v.hex()] for k, v in kvs], root=roots[t].tobytes().hex(), proofs=[])
        for k in targets:
            _, nodes = o.trie_get_proof(kv, t, k)
            assert nodes == T.proof(k)
            jobs.append((t, len(ent["proofs"]), roots[t].tobytes(), nodes, k))
            ent["proofs"].append(dict(key=k.hex(), nodes=[n.hex() for n in nodes]))
        out_tries.append(ent)

    def run(j):
        return ref.run(j[2], j[3], j[4])

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(run, jobs))
    n_ok = 0
    for (t, pi, root, nodes, k), r in zip(jobs, res):
        p = out_tries[t]["proofs"][pi]
        p["status"] = r["status"]
        p["value"] = None if r["value"] is None else r["value"].hex()
        final = dict(tries[t])
        v = final.get(k, b"")
        if r["status"] == 0:
            assert r["value"] == v or (len(v) == 1 and r["value"] == b"\x81" + v), (t, k.hex())
            n_ok += 1
    import hashlib
    doc = dict(generator="oracle/gen_rebuild_golden.py", reference_elf_sha256=hashlib.sha256(ref.elf).hexdigest(),
               note="status/value of every proof come from the reference ELF; roots from two independent builders",
               tries=out_tries)
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    print(len(out_tries), "tries,", len(jobs), "proofs judged by the reference ELF,", n_ok, "accepted ->", OUT,
          os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
