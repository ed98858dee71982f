"""Differential fuzz of the GPU trie rebuild (roots and sampled get_proof output) against the CPU restatement of
eth_trie's insert / root_hash / get_proof.   python tools/fuzz_rebuild_gpu_vs_oracle.py [n_seeds] [first_seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from oracle.pyoracle import Oracle
from tests.test_rebuild_oracle import make_kv, prefix_heavy_tries, random_tries

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
first = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
o = Oracle()
ver = z.Verifier([0])
total = bad = proofs = 0
t0 = time.time()
for seed in range(first, first + n_seeds):
    for gen_name, tries in (("random", random_tries(seed, 600)), ("prefix", prefix_heavy_tries(seed, 1200))):
        d = make_kv(tries)
        kv = z.KvBatch(d["key_bytes"], d["key_off"], d["value_bytes"], d["value_off"], d["value_len"], d["trie_first"])
        want = o.trie_roots(d, nthreads=os.cpu_count() or 1)[0]
        for fused in (1, 0):
            ver.set_option("fused_leaf_hash", fused)
            targets = [(t, kvs[len(kvs) // 2][0]) for t, kvs in enumerate(tries) if kvs][:200]
            roots, b = ver.trie_proofs(kv, targets)
            diff = np.nonzero((roots != want).any(axis=1))[0]
            bad += len(diff)
            for t in diff[:3]:
                print("ROOT MISMATCH", seed, gen_name, fused, int(t), len(tries[t]))
            for q, (t, k) in enumerate(targets):
                got = [b.node_bytes[int(b.node_off[i]):int(b.node_off[i]) + int(b.node_len[i])].tobytes()
                       for i in range(int(b.proof_first[q]), int(b.proof_first[q + 1]))]
                if got != o.trie_get_proof(d, t, k)[1]:
                    bad += 1
                    print("PROOF MISMATCH", seed, gen_name, fused, t, k.hex())
            proofs += len(targets)
        total += len(tries)
    print(f"seed {seed}: cumulative {total} tries (x2 leaf modes), {proofs} proofs, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
print(f"TOTAL {total} tries x 2 leaf modes, {proofs} proofs compared node for node: {bad} mismatches")
