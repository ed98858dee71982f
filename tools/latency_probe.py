"""Where the time of ONE verify_merkle_proof call goes (config 1 proof): the raw C-ABI call with prebuilt ctypes
structures (no numpy / wrapper overhead) beside the Python mirror's call, latency path on and off."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import zk_state_proofs_b200 as z
from zk_state_proofs_b200 import crypto_ops as co
root, proof, key, want = bench.config1_input()
ver = z.Verifier([0])
b = z.flatten([z.MerkleProofInput(proof, root, key)])
st = np.zeros(1, np.uint8); vo = np.zeros(1, np.uint64); vl = np.zeros(1, np.uint32)
cb = co._CBatch(b.node_bytes.ctypes.data, len(b.node_bytes), b.node_off.ctypes.data, b.node_len.ctypes.data, b.n_nodes,
                b.proof_first.ctypes.data, 1, b.roots.ctypes.data, b.key_bytes.ctypes.data, b.key_off.ctypes.data, None)
cr = co._CResult(st.ctypes.data, vo.ctypes.data, vl.ctypes.data)
L = ver.lib
for lp in (1, 0):
    ver.set_option("latency_path", lp)
    for _ in range(500):
        L.mptv_verify_batch(ver.ctx, ctypes.byref(cb), ctypes.byref(cr))
    n = 5000
    t0 = time.perf_counter()
    for _ in range(n):
        L.mptv_verify_batch(ver.ctx, ctypes.byref(cb), ctypes.byref(cr))
    raw = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        ver.verify_batch(b)
    wrap = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        ver.verify_merkle_proof(root, proof, key)
    full = (time.perf_counter() - t0) / n
    assert st[0] == 0 and b.value(int(vo[0]), int(vl[0])) == want
    print(f"latency_path={lp}: raw C-ABI call {raw * 1e6:6.1f} us | Verifier.verify_batch {wrap * 1e6:6.1f} us | verify_merkle_proof (flatten + call + slice) {full * 1e6:6.1f} us", flush=True)
print("nodes", b.node_len.tolist(), "keccak-f", b.n_perm())
if hasattr(L, "mptv_debug_small_clocks"):
    ver.set_option("latency_path", 1)
    L.mptv_verify_batch(ver.ctx, ctypes.byref(cb), ctypes.byref(cr))
    ck = (ctypes.c_longlong * 5)()
    L.mptv_debug_small_clocks(ver.ctx, ck)
    c = list(ck)
    print("latency kernel phases (SM clocks @ 1.965 GHz -> us): load %.1f | hash + decode %.1f | walk %.1f | store + fence %.1f | total %.1f" % (
        c[1] / 1965, (c[2] - c[1]) / 1965, (c[3] - c[2]) / 1965, (c[4] - c[3]) / 1965, c[4] / 1965))
