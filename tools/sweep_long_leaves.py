"""Sweep of the rebuild's long-leaf launch (threshold bin, CTAs per SM) on config 4.   python tools/sweep_long_leaves.py [blocks]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from workload import gen

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
kv = gen.block_tries(blocks, 300, "both", seed=4)
ver = z.Verifier([0])
dev = torch.device("cuda", 0)
to_dev = lambda x: torch.from_numpy(x.view(np.uint8) if x.dtype != np.uint8 else x).to(dev)
d_in = {k: to_dev(getattr(kv, k)) for k in ["key_bytes", "key_off", "value_bytes", "value_off", "value_len", "trie_first"]}
d_roots = torch.zeros(32 * kv.n_tries, dtype=torch.uint8, device=dev)
ptrs = {k: v.data_ptr() for k, v in d_in.items()}
ref = None
for long_bin in (128, 65, 49, 33, 17):
    for ctas in (1, 2, 3):
        if long_bin == 128 and ctas > 1:
            continue
        ver.set_option("long_leaf_bin", long_bin)
        ver.set_option("long_leaf_ctas", ctas)
        ts = []
        for it in range(5):
            ver.trie_roots_device(0, ptrs, kv.n_items, kv.n_tries, d_roots.data_ptr(), value_bytes_len=len(kv.value_bytes))
            t = ver.last_rebuild_timings(0)
            if it >= 2:
                ts.append((t.keccak_ms, t.total_ms))
        r = d_roots.cpu().numpy().copy()
        if ref is None:
            ref = r
        assert (r == ref).all()
        k, tot = np.mean(ts, axis=0)
        print(f"blocks={blocks} long_bin>={long_bin:3d} ctas/SM={ctas}: keccak {k:7.3f} ms  total {tot:7.3f} ms")
