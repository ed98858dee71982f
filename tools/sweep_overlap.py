"""Sweep of the overlapped device pipeline ("overlap_ranges" x "overlap_keccak_ctas") on one workload, built once.
usage (on a B200): python tools/sweep_overlap.py [config2|config3] [steps]   -> one line per setting"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zk_state_proofs_b200 as z  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sys.argv = ["bench.py", "--workload", workload]
a = bench.parse_args()
dev = torch.device("cuda", 0)
ver = z.Verifier([0])
b, _ = bench.build_batch(a, 0, pinned=False)
n_proofs, n_nodes = b.n_proofs, b.n_nodes


def to_dev(x):
    return torch.from_numpy(x.view(np.uint8) if x.dtype != np.uint8 else x).to(dev)


names = ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_bytes", "key_off"]
if b.root_from_proof is not None:
    names.append("root_from_proof")
d_in = {k: to_dev(getattr(b, k)) for k in names}
d_status = torch.zeros(n_proofs, dtype=torch.uint8, device=dev)
d_voff = torch.zeros(n_proofs, dtype=torch.int64, device=dev)
d_vlen = torch.zeros(n_proofs, dtype=torch.int32, device=dev)
ptrs = {k: v.data_ptr() for k, v in d_in.items()}
outp = dict(status=d_status.data_ptr(), value_off=d_voff.data_ptr(), value_len=d_vlen.data_ptr())
stream = torch.cuda.Stream(device=dev)
ref = None
ver.set_option("overlap_min_nodes", 1)
for S, C in [(0, 3), (2, 3), (3, 3), (4, 3), (6, 3), (8, 3), (12, 3), (16, 3), (4, 4), (8, 4), (4, 2), (8, 2), (0, 3)]:
    ver.set_option("overlap_ranges", S)
    ver.set_option("overlap_keccak_ctas", C)
    for _ in range(3):
        ver.verify_batch_device(0, ptrs, n_nodes, n_proofs, outp, stream=stream.cuda_stream, node_bytes_len=len(b.node_bytes))
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        ver.verify_batch_device(0, ptrs, n_nodes, n_proofs, outp, stream=stream.cuda_stream, node_bytes_len=len(b.node_bytes))
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    t = ver.last_timings(0)
    res = (d_status.cpu().numpy().copy(), d_voff.cpu().numpy().copy(), d_vlen.cpu().numpy().copy())
    if ref is None:
        ref = res
    same = all((x == y).all() for x, y in zip(ref, res))
    print(f"ranges {S:2d} k1_ctas {C}: {ms:7.3f} ms/pass  {n_proofs / ms / 1e3:7.2f} M proofs/s  "
          f"bin {t.bin_ms:.3f} keccak {t.keccak_ms:.3f} parse {t.parse_ms:.3f} walk/exposed {t.walk_ms:.3f}  "
          f"results identical to sequential: {same}", flush=True)
