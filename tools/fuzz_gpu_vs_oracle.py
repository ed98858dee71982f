"""Large differential fuzz of the CUDA path against the C restatement (which is itself fuzzed against the reference
ELF by oracle/fuzz_vs_ref.py): every fuzz family of oracle/fuzzgen.py, all walk configurations.
    python tools/fuzz_gpu_vs_oracle.py [n_seeds] [first_seed]"""
import os
import sys
import time
from collections import Counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from oracle.fuzzgen import corpus
from oracle.pyoracle import Oracle

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
first = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
o = Oracle()
ver = z.Verifier([0])
modes = [dict(fast_walk=1, dedup_nodes=0), dict(fast_walk=0, dedup_nodes=0), dict(fast_walk=1, dedup_nodes=1)]
total, bad, hist = 0, 0, Counter()
t0 = time.time()
for seed in range(first, first + n_seeds):
    cases = corpus(seed, o.keccak256, 300, 15000, 20000, 15000, 30000, 4000)
    b = z.flatten_borsh([z.MerkleProofInput(c["proof"], c["root"], c["key"]).to_borsh() for c in cases], threads=0)
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off)
    ost, ovoff, ovlen, _, _ = o.verify_batch(d, nthreads=os.cpu_count() or 1)
    if b.bad_root_len is not None:
        ost = ost.copy(); ost[b.bad_root_len] = 6
        ovoff = ovoff.copy(); ovoff[b.bad_root_len] = 0
        ovlen = ovlen.copy(); ovlen[b.bad_root_len] = 0
    for m in modes:
        for k, v in m.items():
            ver.set_option(k, v)
        st, voff, vlen = ver.verify_batch(b)
        diff = np.nonzero((st != ost) | (voff != ovoff) | (vlen != ovlen))[0]
        bad += len(diff)
        for i in diff[:5]:
            print("MISMATCH", seed, m, cases[i]["tag"], int(st[i]), int(ost[i]))
    total += len(cases)
    hist.update(ost.tolist())
    print(f"seed {seed}: {len(cases)} cases x {len(modes)} modes, cumulative mismatches {bad}, {time.time() - t0:.0f} s", flush=True)
print(f"TOTAL {total} cases x {len(modes)} walk configurations: {bad} mismatches; verdict histogram {dict(sorted(hist.items()))}")
