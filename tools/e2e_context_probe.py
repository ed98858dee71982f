"""Why does the borsh e2e leg inside bench.py run slower than the same call in a bare script?  Times the same
mptv_verify_borsh call (mode 0, 14 threads) in a fresh process with (a) nothing else, (b) torch imported and a CUDA
tensor copy done first, (c) after device-resident verify steps, printing every iteration."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import zk_state_proofs_b200 as z  # noqa: E402
from workload import gen  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "bare"
sys.argv = ["bench.py", "--workload", "config2"]
a = bench.parse_args()
ver = z.Verifier([0])
b, _ = bench.build_batch(a, 0, pinned=True)
th = (os.cpu_count() or 16) - 2
if what in ("torch", "omp1"):
    import torch
    if what == "omp1":
        torch.set_num_threads(1)
    x = torch.from_numpy(b.node_bytes).to("cuda:0")
    y = x[:1000000].cpu().numpy().sum()
    torch.cuda.synchronize()
if what == "csr":
    for _ in range(3):
        ver.verify_batch(b)
blobs, boff = gen.batch_to_borsh(b, pinned=True)
ts = []
for it in range(8):
    ver.host_stats(reset=True)
    t0 = time.perf_counter()
    ver.verify_borsh(blobs, boff, threads=th)
    ts.append((time.perf_counter() - t0) * 1e3)
hs = ver.host_stats()
print(f"{what:6s}: " + " ".join(f"{t:5.1f}" for t in ts) + f" ms | last: flatten {hs.flatten_us / 1e3:.1f} wait {hs.wait_us / 1e3:.1f}", flush=True)
