"""tools/sweep_keccak.py -- time K1 (k_keccak256_nodes) alone on a synthetic arena shaped like
config-2 proofs (node length histogram of a 10 M-account trie).  Run once per libmptv variant:
    MPTV_LIB=/path/to/libmptv_variant.so python tools/sweep_keccak.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z

I_PERM = 4320


def main():
    n = int(os.environ.get("SWEEP_NODES", 4_000_000))
    rng = np.random.default_rng(0)
    # per-proof node lengths ~ [532 x5, ~400, ~150, ~105]
    choices = np.array([532, 532, 532, 532, 532, 404, 147, 104], np.uint32)
    lens = choices[rng.integers(0, len(choices), n)]
    lens = (lens + rng.integers(0, 3, n).astype(np.uint32) * (lens < 500)).astype(np.uint32)
    padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
    off = np.zeros(n, np.uint64)
    np.cumsum(padded[:-1], out=off[1:])
    total = int(padded.sum()) + 16
    dev = torch.device("cuda", 0)
    nb = torch.randint(0, 256, (total,), dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_len = torch.from_numpy(lens.view(np.int32)).to(dev)
    dig = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    ver = z.Verifier([0])
    n_perm = int((lens // 136 + 1).sum())
    peak = max(ver.int_issue_peak(0, m) for m in (0, 2))
    best = {}
    for binning in (1, 0):
        ver.set_option("binning", binning)
        ts = []
        for it in range(8):
            ver.keccak256_batch_device(0, nb.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), n, dig.data_ptr())
            t = ver.last_timings(0)
            if it >= 3:
                ts.append((t.keccak_ms, t.bin_ms))
        k = float(np.mean([a for a, _ in ts]))
        b = float(np.mean([c for _, c in ts]))
        best[binning] = k
        print(f"lib={os.path.basename(z.lib_path())} binning={binning} keccak_ms={k:.3f} bin_ms={b:.3f} "
              f"Gkeccak/s={n_perm / k / 1e6:.3f} frac_of_int_peak={n_perm * I_PERM / (k * 1e-3) / peak:.4f} "
              f"peak_Tops={peak / 1e12:.2f}")
    # spot check vs the first digests computed on the CPU is done by the parity tests, not here
    h = torch.sum(dig.view(torch.int32).to(torch.int64)).item()
    print("digest checksum", h)


if __name__ == "__main__":
    main()
