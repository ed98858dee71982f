"""Repeated calls with varying batch sizes and options: results stay equal to the oracle, device memory does not grow
without bound.   python tools/stress_sizes.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from oracle.pyoracle import Oracle
from workload import gen
from zk_state_proofs_b200.sharding import take_slice

o = Oracle()
ver = z.Verifier([0])
state, tokens = gen.make_state_and_tokens(200_000, 3, 20_000, seed=3)
big = gen.nested_batch(state, tokens, 30_000, seed=21)
d = dict(node_bytes=big.node_bytes, node_off=big.node_off, node_len=big.node_len, proof_first=big.proof_first,
         roots=big.roots, key_bytes=big.key_bytes, key_off=big.key_off, root_from_proof=big.root_from_proof)
ost, ovoff, ovlen, _, _ = o.verify_batch(d, nthreads=8)
kv = gen.block_tries(300, 300, "both", seed=4)
want_roots = o.trie_roots(kv.as_dict(), nthreads=8)[0]
rng = np.random.default_rng(0)
free0 = None
for it in range(120):
    g0 = int(rng.integers(0, 29_000))
    g1 = min(30_000, g0 + int(rng.choice([1, 2, 7, 100, 1000, 5000])))
    s = take_slice(big, 4 * g0, 4 * g1)
    ver.set_option("chunk_bytes", int(rng.choice([1 << 16, 1 << 20, 96 << 20])))
    ver.set_option("fast_walk", int(rng.integers(0, 2)))
    ver.set_option("dedup_nodes", int(rng.integers(0, 2)))
    st, voff, vlen = ver.verify_batch(s)
    assert (st == ost[4 * g0:4 * g1]).all() and (vlen == ovlen[4 * g0:4 * g1]).all(), it
    if it % 10 == 0:
        t0 = int(rng.integers(0, 500))
        t1 = min(600, t0 + int(rng.choice([1, 3, 50, 100])))
        ni0, ni1 = int(kv.trie_first[t0]), int(kv.trie_first[t1])
        sub = z.KvBatch(kv.key_bytes, kv.key_off[ni0:ni1 + 1] if False else kv.key_off, kv.value_bytes, kv.value_off, kv.value_len, kv.trie_first[:t1 + 1])
        assert (ver.trie_roots(sub) == want_roots[:t1]).all()
    free, total = torch.cuda.mem_get_info()
    if it == 20:
        free0 = free
print("ok; device memory used after warm-up grew by", (free0 - torch.cuda.mem_get_info()[0]) / 1e6, "MB over 100 further calls")
