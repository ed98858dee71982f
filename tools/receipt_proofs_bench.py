"""Closed loop at scale: rebuild receipt tries on the GPU, extract the inclusion proof of EVERY receipt
(mptv_trie_proofs), verify them all (device-resident) and check the returned values.  Shows K1 on proofs
whose leaves are long (receipts up to 30 KB = 221 rate blocks).   python tools/receipt_proofs_bench.py [blocks]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
from workload import gen

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
kv = gen.block_tries(blocks, 300, "receipt", seed=4)
ver = z.Verifier([0])
keys = [z.rlp_index(i) for i in range(300)]
targets = [(t, keys[i]) for t in range(blocks) for i in range(300)]
t0 = time.time()
roots, b = ver.trie_proofs(kv, targets)
print(f"rebuild + get_proof of {len(targets)} targets: {time.time() - t0:.2f} s (host call incl. Python marshalling); "
      f"{b.n_nodes} nodes, {b.node_len.astype(np.int64).sum() / 1e9:.2f} GB, {b.n_perm() / 1e6:.1f} M Keccak-f")
dev = torch.device("cuda", 0)
t = {k: torch.from_numpy(getattr(b, k).view(np.uint8) if getattr(b, k).dtype != np.uint8 else getattr(b, k)).to(dev)
     for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_bytes", "key_off"]}
d_st = torch.zeros(b.n_proofs, dtype=torch.uint8, device=dev)
d_vo = torch.zeros(b.n_proofs, dtype=torch.int64, device=dev)
d_vl = torch.zeros(b.n_proofs, dtype=torch.int32, device=dev)
ptrs = {k: v.data_ptr() for k, v in t.items()}
outp = dict(status=d_st.data_ptr(), value_off=d_vo.data_ptr(), value_len=d_vl.data_ptr())
for _ in range(4):
    ver.verify_batch_device(0, ptrs, b.n_nodes, b.n_proofs, outp, node_bytes_len=len(b.node_bytes))
    tm = ver.last_timings(0)
st = d_st.cpu().numpy()
vl = d_vl.cpu().numpy().view(np.uint32)
assert (st == 0).all() and (vl == kv.value_len).all()
int_peak = max(ver.int_issue_peak(0, m) for m in (0, 2))
print(f"verify: total {tm.total_ms:.3f} ms  keccak {tm.keccak_ms:.3f} ms  walk {tm.walk_ms:.3f} ms -> "
      f"{b.n_proofs / tm.total_ms / 1e3:.1f} M proofs/s, K1 at {b.n_perm() * 4320 / (tm.keccak_ms * 1e-3) / int_peak:.3f} of the integer-issue peak")
