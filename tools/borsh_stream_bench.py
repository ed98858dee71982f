"""mptv_verify_borsh (blobs in, verdicts out) on the config-2 batch: host_dedup x chunk size x host threads sweep,
with the bytes that crossed PCIe (mptv_host_stats) and the host stage alone (mptv_borsh_flatten_probe) beside it,
and the serial flatten-then-verify.   python tools/borsh_stream_bench.py [n_proofs] [quick]      (on a B200)"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zk_state_proofs_b200 as z  # noqa: E402
from workload import gen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
quick = len(sys.argv) > 2
pinned_blobs = len(sys.argv) > 3  # page-locked blobs: the entry runs in pull mode
sys.argv = ["bench.py", "--workload", "config2", "--proofs", str(n)]
a = bench.parse_args()
ver = z.Verifier([0])
if len(sys.argv) > 3:
    ver.set_option("pull_pinned", 1)
b, _ = bench.build_batch(a, 0, pinned=True)
blobs, boff = gen.batch_to_borsh(b, pinned=pinned_blobs)
cores = os.cpu_count()
print(f"{n} proofs, {len(blobs) / 1e9:.2f} GB of borsh ({'page-locked: pull mode' if pinned_blobs else 'pageable'}), {cores} host cores", flush=True)
ref = ver.verify_batch(b)
for _ in range(2):
    t0 = time.perf_counter(); ver.verify_batch(b); dt = time.perf_counter() - t0
print(f"mptv_verify_batch on the pre-flattened pinned batch: {dt * 1e3:.1f} ms = {n / dt / 1e6:.2f} M proofs/s", flush=True)
for dd in (1, 0):
    ver.set_option("host_dedup", dd)
    for mb in ((32,) if quick else (16, 32, 64)):
        for th in ((0,) if quick else sorted({4, 8, 12, cores - 1, cores})):  # 0 = the library's default (all cores but one)
            ver.set_option("borsh_chunk_bytes", mb << 20)
            best = 1e9
            for it in range(4):
                ver.host_stats(reset=True)
                t0 = time.perf_counter()
                st, voff, vlen = ver.verify_borsh(blobs, boff, threads=th)
                dt = time.perf_counter() - t0
                if it:
                    best = min(best, dt)
            hs = ver.host_stats()
            assert (st == ref[0]).all() and (vlen == ref[2]).all()
            pt = min(z.borsh_flatten_probe(blobs, boff, threads=th, chunk_bytes=mb << 20, alias_duplicates=bool(dd))[0]
                     for _ in range(3))
            print(f"host_dedup {dd} chunk {mb:4d} MB threads {th or cores - 1:2d}: {best * 1e3:7.1f} ms = {n / best / 1e6:6.2f} M proofs/s | "
                  f"H2D {hs.h2d_bytes / 1e9:5.2f} GB ({hs.h2d_bytes / best / 1e9:5.1f} GB/s), aliased {hs.nodes_aliased}/{hs.nodes} nodes, "
                  f"{hs.chunks} chunks ({hs.pull_chunks} pulled), flatten {hs.flatten_us / 1e3:.1f} ms | host stage alone {pt * 1e3:6.1f} ms ({len(blobs) / pt / 1e9:5.1f} GB/s read)", flush=True)
ver.set_option("borsh_chunk_bytes", 32 << 20)
ver.set_option("host_dedup", 1)
# device flatten: page-locked blobs cross PCIe as they are, kernels lay the nodes out
pblobs, pboff = gen.batch_to_borsh(b, pinned=True)
ver.set_option("borsh_mode", 1)
for mb in (16, 32, 64):
    ver.set_option("borsh_chunk_bytes", mb << 20)
    best = 1e9
    for it in range(4):
        ver.host_stats(reset=True)
        t0 = time.perf_counter()
        dst, dvoff, dvlen = ver.verify_borsh(pblobs, pboff)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    hs = ver.host_stats()
    assert (dst == ref[0]).all() and (dvlen == ref[2]).all() and (dvoff == voff).all()
    print(f"device flatten chunk {mb:4d} MB: {best * 1e3:7.1f} ms = {n / best / 1e6:6.2f} M proofs/s | H2D {hs.h2d_bytes / 1e9:5.2f} GB "
          f"({hs.h2d_bytes / best / 1e9:5.1f} GB/s), {hs.device_chunks} chunks on the device", flush=True)
# hybrid: both pipelines on the one device, chunks from the two ends of the input
ver.set_option("borsh_mode", 2)
for mb, pct in ((32, 15), (32, 20), (32, 24), (32, 30), (32, 40), (16, 24), (64, 24)):
    for th in (0,):
        ver.set_option("borsh_chunk_bytes", mb << 20)
        ver.set_option("hybrid_device_pct", pct)
        best = 1e9
        for it in range(5):
            ver.host_stats(reset=True)
            t0 = time.perf_counter()
            dst, dvoff, dvlen = ver.verify_borsh(pblobs, pboff, threads=th)
            dt = time.perf_counter() - t0
            if it:
                best = min(best, dt)
        hs = ver.host_stats()
        assert (dst == ref[0]).all() and (dvlen == ref[2]).all() and (dvoff == voff).all()
        print(f"hybrid {pct:2d} % chunk {mb:4d} MB threads {th or cores - 1:2d}: {best * 1e3:7.1f} ms = {n / best / 1e6:6.2f} M proofs/s | H2D {hs.h2d_bytes / 1e9:5.2f} GB "
              f"({hs.h2d_bytes / best / 1e9:5.1f} GB/s), {hs.device_chunks} of {hs.chunks} chunks on the device, aliased {hs.nodes_aliased}", flush=True)
ver.set_option("borsh_mode", 0)
ver.set_option("borsh_chunk_bytes", 32 << 20)
ok = np.nonzero(st == 0)[0][:5000]
for i in ok:
    assert blobs[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() == b.value(int(ref[1][i]), int(ref[2][i]))
L = z.load_library()
h = ctypes.c_void_p()
for it in range(3):
    t0 = time.perf_counter()
    assert L.mptv_flatten_borsh(blobs.ctypes.data, boff.ctypes.data, n, 0, 1, ctypes.byref(h)) == 0
    t1 = time.perf_counter()
    ver.verify_batch(z.crypto_ops.batch_from_handle(L, h, n))
    t2 = time.perf_counter()
print(f"serial: flatten {1e3 * (t1 - t0):.1f} ms + verify_batch {1e3 * (t2 - t1):.1f} ms = {n / (t2 - t0) / 1e6:.2f} M proofs/s")
L.mptv_host_batch_free(h)
