"""mptv_verify_borsh (mode 0) with the node bytes staged in ordinary vs write-combining page-locked memory.
    python tools/wc_probe.py [n_proofs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import zk_state_proofs_b200 as z  # noqa: E402
from workload import gen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sys.argv = ["bench.py", "--workload", "config2", "--proofs", str(n)]
a = bench.parse_args()
ver = z.Verifier([0])
b, _ = bench.build_batch(a, 0, pinned=True)
ref = ver.verify_batch(b)
cores = os.cpu_count()
for pinned in (True, False):
    blobs, boff = gen.batch_to_borsh(b, pinned=pinned)
    for rep in range(2):
        for wc in (0, 1):
            for th in (cores - 2, cores - 1):
                ver.set_option("wc_staging", wc)
                ts = []
                for it in range(6):
                    ver.host_stats(reset=True)
                    t0 = time.perf_counter()
                    st, voff, vlen = ver.verify_borsh(blobs, boff, threads=th)
                    ts.append((time.perf_counter() - t0) * 1e3)
                hs = ver.host_stats()
                assert (st == ref[0]).all() and (vlen == ref[2]).all()
                print(f"blobs {'page-locked' if pinned else 'pageable   '} wc_staging {wc} threads {th}: best {min(ts[1:]):5.1f} mean {sum(ts[1:]) / 5:5.1f} ms | "
                      f"last: flatten {hs.flatten_us / 1e3:5.1f} wait {hs.wait_us / 1e3:4.1f} map {hs.map_us / 1e3:4.1f}", flush=True)
ver.set_option("wc_staging", 0)
