"""host stage of mptv_verify_borsh alone (mptv_borsh_flatten_probe) for each library given in MPTV_LIBS (':' separated)
   python tools/flatten_probe_sweep.py [n_proofs]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("MPTV_LIB") is None or len(sys.argv) > 2:
    pass
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import zk_state_proofs_b200 as z
from workload import gen
import numpy as np
blobs = np.fromfile("/tmp/blobs.bin", np.uint8); boff = np.fromfile("/tmp/boff.bin", np.uint64)
n = len(boff) - 1
for th in (8, 12, 15, 16):
    best = min(z.borsh_flatten_probe(blobs, boff, threads=th, chunk_bytes=32 << 20)[0] for _ in range(6))
    print(f"{os.path.basename(os.environ.get('MPTV_LIB', 'libmptv.so')):28s} threads {th:2d}: {best * 1e3:7.1f} ms  {len(blobs) / best / 1e9:6.1f} GB/s read  {n / best / 1e6:6.2f} M proofs/s", flush=True)
''' % ROOT
libs = [os.path.join(ROOT, "zk-state-proofs_b200", "libmptv.so")] + sorted(
    os.path.join(ROOT, "build", "variants", f) for f in os.listdir(os.path.join(ROOT, "build", "variants")) if f.startswith("libmptv_ahead"))
for lib in libs:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MPTV_LIB=lib))
