"""Where the hybrid borsh mode (borsh_mode 2) loses its time: per setting, the host pipeline's own clocks
(mptv_host_stats: flatten / wait / map) beside the call's wall time.   python tools/hybrid_probe.py [n_proofs]"""
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
blobs, boff = gen.batch_to_borsh(b, pinned=True)
ref = ver.verify_batch(b)
cores = os.cpu_count()


def run(label, mode, th, mb=32, pct=24):
    ver.set_option("borsh_mode", mode)
    ver.set_option("borsh_chunk_bytes", mb << 20)
    ver.set_option("hybrid_device_pct", pct)
    best, hb = 1e9, None
    for it in range(5):
        ver.host_stats(reset=True)
        t0 = time.perf_counter()
        st, voff, vlen = ver.verify_borsh(blobs, boff, threads=th)
        dt = time.perf_counter() - t0
        if it and dt < best:
            best, hb = dt, ver.host_stats()
    assert (st == ref[0]).all() and (vlen == ref[2]).all()
    print(f"{label:34s} threads {th:2d} chunk {mb:3d} MB: {best * 1e3:6.1f} ms = {n / best / 1e6:5.2f} M/s | host pipeline: flatten "
          f"{hb.flatten_us / 1e3:5.1f} wait {hb.wait_us / 1e3:5.1f} map {hb.map_us / 1e3:4.1f} call {hb.call_us / 1e3:5.1f} ms | "
          f"{hb.device_chunks}/{hb.chunks} chunks on the device, H2D {hb.h2d_bytes / 1e9:4.2f} GB", flush=True)


print(f"{n} proofs, {len(blobs) / 1e9:.2f} GB of page-locked borsh, {cores} cores", flush=True)
for th in (cores - 1, cores - 2):
    run("host flatten (mode 0)", 0, th)
run("device flatten (mode 1)", 1, 0)
for piece in (8 << 20, 0, 2 << 20):
    ver.set_option("hybrid_copy_piece", piece)
    for pct in (20, 24, 30, 40):
        for th in (cores - 2, cores - 3):
            run(f"hybrid {pct} %, pieces of {piece >> 20} MB", 2, th, 32, pct)
ver.set_option("hybrid_copy_piece", 8 << 20)
ver.set_option("borsh_mode", 0)
