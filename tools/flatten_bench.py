"""Throughput of the multi-threaded borsh(MerkleProofInput) -> CSR flattener (csrc/host_codec.cpp), the
host step right in front of mptv_verify_batch.  CPU only.   python tools/flatten_bench.py [n_proofs] [threads...]"""
import ctypes
import os
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from workload import gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
threads = [int(x) for x in sys.argv[2:]] or [1, 4, 0]
trie = gen.SynthTrie(2_000_000, 2, kind=0)
b = gen.account_batch(trie, n, seed=2)
# serialise the batch as borsh blobs with numpy (bulk), as a prover's input file would hold them
blobs = bytearray()
off = [0]
for p in range(n):
    a, e = int(b.proof_first[p]), int(b.proof_first[p + 1])
    blobs += struct.pack("<I", e - a)
    for i in range(a, e):
        o, ln = int(b.node_off[i]), int(b.node_len[i])
        blobs += struct.pack("<I", ln) + b.node_bytes[o:o + ln].tobytes()
    blobs += struct.pack("<I", 32) + b.roots[32 * p:32 * p + 32].tobytes()
    blobs += struct.pack("<I", 32) + b.key_bytes[32 * p:32 * p + 32].tobytes()
    off.append(len(blobs))
buf = np.frombuffer(bytes(blobs), np.uint8)
offs = np.array(off, np.uint64)
L = z.load_library()
for t in threads:
    best = 1e9
    h = ctypes.c_void_p()
    for it in range(4):  # the handle is recycled: the first call pays for the page faults of fresh buffers
        t0 = time.perf_counter()
        rc = L.mptv_flatten_borsh(buf.ctypes.data, offs.ctypes.data, n, t, 0, ctypes.byref(h))
        dt = time.perf_counter() - t0
        assert rc == 0
        if it == 0:
            first = dt
        else:
            best = min(best, dt)
    L.mptv_host_batch_free(h)
    print(f"threads={t or os.cpu_count()}: first call {n / first / 1e6:.2f} M proofs/s; steady state {n / best / 1e6:.2f} M proofs/s, "
          f"{len(buf) / best / 1e9:.2f} GB/s of borsh input")
