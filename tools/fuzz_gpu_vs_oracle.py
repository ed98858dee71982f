"""Large differential fuzz of the CUDA path against the C restatement (which is itself fuzzed against the reference
ELF by oracle/fuzz_vs_ref.py): every fuzz family of oracle/fuzzgen.py incl. the deep inline nests, all walk
configurations, the streamed borsh entry with transfer de-duplication, its device-flatten and hybrid modes on
page-locked blobs, and the one-launch latency path on random small groups.
    python tools/fuzz_gpu_vs_oracle.py [n_seeds] [first_seed]"""
import os
import sys
import time
from collections import Counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
import random

from oracle.fuzzgen import corpus, deep_nested_cases
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
    cases += deep_nested_cases(random.Random(seed), o.keccak256, 45)
    blobs = [z.MerkleProofInput(c["proof"], c["root"], c["key"]).to_borsh() for c in cases]
    b = z.flatten_borsh(blobs, threads=0)
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
    for k, v in modes[0].items():
        ver.set_option(k, v)
    # the streamed borsh entry, byte-identical nodes of a chunk aliased (values come back as slices of the blobs)
    boff = np.zeros(len(blobs) + 1, np.uint64)
    np.cumsum([len(x) for x in blobs], out=boff[1:])
    buf = np.frombuffer(b"".join(blobs) + b"\0", np.uint8)
    ver.set_option("borsh_chunk_bytes", 1 << 20)
    st, voff, vlen = ver.verify_borsh(buf, boff)
    ver.set_option("borsh_chunk_bytes", 32 << 20)
    diff = np.nonzero((st != ost) | (vlen != ovlen))[0]
    for i in np.nonzero(st == 0)[0][::7]:
        if buf[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() != b.value(int(ovoff[i]), int(ovlen[i])):
            diff = np.append(diff, i)
    bad += len(diff)
    for i in diff[:5]:
        print("MISMATCH borsh stream", seed, cases[i]["tag"], int(st[i]), int(ost[i]))
    # the same blobs in page-locked memory, flattened on the device (borsh_mode 1) and by both pipelines at once (2):
    # verdicts, value offsets and lengths must equal the host-flatten mode's exactly
    import torch
    pin = torch.empty(len(buf) + 64, dtype=torch.uint8, pin_memory=True)
    lead = seed % 16  # any alignment of the blob image
    pbuf = pin.numpy()[lead:lead + len(buf)]
    pbuf[:] = buf
    for mode, chunk in ((1, 1 << 19), (2, 1 << 19), (1, 32 << 20)):
        ver.set_option("borsh_mode", mode)
        ver.set_option("borsh_chunk_bytes", chunk)
        st1, voff1, vlen1 = ver.verify_borsh(pbuf, boff)
        d1 = np.nonzero((st1 != st) | (voff1 != voff) | (vlen1 != vlen))[0]
        bad += len(d1)
        for i in d1[:5]:
            print("MISMATCH borsh mode", mode, seed, cases[i]["tag"], int(st1[i]), int(st[i]))
    ver.set_option("borsh_mode", 0)
    ver.set_option("borsh_chunk_bytes", 32 << 20)
    # the one-launch latency path: random groups of 1 ... 32 proofs per call
    rng = random.Random(seed)
    for _ in range(1500):
        k = rng.choice([1, 1, 1, 2, 4, 9, 32])
        idx = [rng.randrange(len(cases)) for _ in range(k)]
        sb = z.flatten([z.MerkleProofInput(cases[i]["proof"], cases[i]["root"], cases[i]["key"]) for i in idx])
        st, voff, vlen = ver.verify_batch(sb)
        for j, i in enumerate(idx):
            want_v = b.value(int(ovoff[i]), int(ovlen[i])) if ost[i] == 0 else b""
            got_v = sb.value(int(voff[j]), int(vlen[j])) if st[j] == 0 else b""
            if int(st[j]) != int(ost[i]) or got_v != want_v:
                bad += 1
                print("MISMATCH latency path", seed, cases[i]["tag"], int(st[j]), int(ost[i]))
    total += len(cases)
    hist.update(ost.tolist())
    print(f"seed {seed}: {len(cases)} cases x {len(modes)} modes, cumulative mismatches {bad}, {time.time() - t0:.0f} s", flush=True)
print(f"TOTAL {total} cases x ({len(modes)} walk configurations + borsh stream with aliasing + device-flatten and hybrid borsh modes) + 1500 latency-path groups per seed: {bad} mismatches; "
      f"verdict histogram {dict(sorted(hist.items()))}")
