"""Differential campaign for mptv_verify_storage_borsh: random StorageProofInputs (tests/test_gpu_storage_borsh.py's
generator: real account leaves, leaves that are no Account, unequal proof / key lists, bad root lengths, mutated /
absent / shuffled proofs) serialised, streamed through the CUDA path, and compared input by input with the storage
guest restated over the C oracle (tests/test_gpu_storage.py::_guest).
    python tools/fuzz_storage_stream.py [seconds] [first_seed]"""
import os
import sys
import time
from collections import Counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from tests.test_gpu_storage_borsh import _guest_flow, _inputs  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
o = Oracle()
ver = z.Verifier([0])
t0 = time.time()
total = bad = proofs = 0
hist = Counter()
while time.time() - t0 < budget:
    inputs = _inputs(o, seed, n_groups=600)
    want = _guest_flow(o, inputs)
    blobs = [i.to_borsh() for i in inputs]
    buf = np.frombuffer(b"".join(blobs), np.uint8)
    for chunk, dd in ((1 << 14, 1), (32 << 20, 1), (1 << 16, 0)):
        ver.set_option("borsh_chunk_bytes", chunk)
        ver.set_option("host_dedup", dd)
        pf, ist, st, voff, vlen = ver.verify_storage_borsh(blobs, threads=4)
        for i, w in enumerate(want):
            a, e = int(pf[i]), int(pf[i + 1])
            if isinstance(w, z.VerifyPanic):
                ok = int(ist[i]) == w.status
            else:
                ok = ist[i] == 0 and [buf[int(voff[q]):int(voff[q]) + int(vlen[q])].tobytes() for q in range(a + 1, e)] == w
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, "input", i, "chunk", chunk, int(ist[i]), w if isinstance(w, z.VerifyPanic) else "values", flush=True)
    proofs += int(pf[-1])
    total += len(inputs)
    hist.update(0 if isinstance(w, list) else w.status for w in want)
    seed += 1
print(f"TOTAL {total} inputs ({proofs} proofs) x 3 pipeline settings: {bad} mismatches; guest outcomes {dict(sorted(hist.items()))}; "
      f"{time.time() - t0:.0f} s")
