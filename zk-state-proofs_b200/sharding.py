"""Multi-process sharding of a batch: one process per GPU, independent proof slices, NO data-path
collective.  The only communication is the final gather of the 13-byte
per-proof results (and, in bench.py, the barrier / max-over-ranks of the timing).

The cut rule is the one libmptv.so applies inside one process to the devices of a context
(mptv_api.cu: mptv_verify_batch): contiguous proof ranges holding equal shares of the node bytes
(proportional to the Keccak-f count), never separating a storage proof from the account proof its
root comes from (`root_from_proof`).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np

from .crypto_ops import Batch


def slice_cuts(b: Batch, world: int) -> List[int]:
    """-> world + 1 proof indices; rank r owns proofs [cuts[r], cuts[r + 1])."""
    n = b.n_proofs
    cuts = [0] * (world + 1)
    cuts[world] = n
    if n == 0:
        return cuts
    n_total = int(b.proof_first[n])
    total = int(b.node_off[n_total - 1]) + int(b.node_len[n_total - 1]) if n_total else 0
    # byte offset at which each proof starts (proofs without nodes start where the next one does)
    first = b.proof_first[:n].astype(np.int64)
    starts = np.where(first < n_total, b.node_off[np.minimum(first, max(n_total - 1, 0))].astype(np.int64), total)
    p = 0
    for k in range(1, world):
        target = total // world * k
        p = max(p, int(np.searchsorted(starts, target, side="left")))
        if b.root_from_proof is not None:
            while p < n and b.root_from_proof[p] >= 0:
                p += 1
        cuts[k] = p
    return cuts


def take_slice(b: Batch, p0: int, p1: int) -> Batch:
    """Proofs [p0, p1) as a self-contained batch (offsets and indices rebased, arrays are views or
    small copies; the node bytes are a view)."""
    n0, n1 = int(b.proof_first[p0]), int(b.proof_first[p1])
    byte0 = int(b.node_off[n0]) if n1 > n0 else 0
    byte1 = (int(b.node_off[n1 - 1]) + int(b.node_len[n1 - 1]) + 15) & ~15 if n1 > n0 else 0
    k0, k1 = int(b.key_off[p0]), int(b.key_off[p1])
    rfp = None
    if b.root_from_proof is not None:
        rfp = b.root_from_proof[p0:p1].copy()
        if ((rfp >= 0) & (rfp < p0)).any():  # checked on the ORIGINAL values: p0 - 1 would rebase to -1 = "independent"
            raise ValueError("slice separates a storage proof from its account proof")
        rfp[rfp >= 0] -= p0
    node_bytes = b.node_bytes[byte0:byte1 + 16]
    if len(node_bytes) < byte1 - byte0 + 16:
        node_bytes = np.concatenate([node_bytes, np.zeros(byte1 - byte0 + 16 - len(node_bytes), np.uint8)])
    return Batch(node_bytes, b.node_off[n0:n1] - np.uint64(byte0), b.node_len[n0:n1],
                 (b.proof_first[p0:p1 + 1] - np.uint32(n0)).astype(np.uint32), b.roots[32 * p0:32 * p1],
                 b.key_bytes[k0:k1 + 16] if k1 + 16 <= len(b.key_bytes) else np.concatenate(
                     [b.key_bytes[k0:k1], np.zeros(16, np.uint8)]),
                 (b.key_off[p0:p1 + 1] - np.uint32(k0)).astype(np.uint32), rfp,
                 None if b.bad_root_len is None else b.bad_root_len[p0:p1])


def verify_sharded(b: Batch, verify_fn: Optional[Callable] = None, group=None):
    """Every rank of the (already initialised) torch.distributed group calls this with the same
    batch; rank r verifies its slice on its own GPU and the per-proof results are all-gathered.
    verify_fn(batch) -> (status u8[n], value_off u64[n], value_len u32[n]); default: a Verifier on
    cuda:LOCAL_RANK.  Returned value offsets refer to the FULL batch's node_bytes."""
    import os
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cuts = slice_cuts(b, world)
    p0, p1 = cuts[rank], cuts[rank + 1]
    if verify_fn is None:
        from .crypto_ops import Verifier
        ver = Verifier([int(os.environ.get("LOCAL_RANK", rank))])
        verify_fn = ver.verify_batch
    if p1 > p0:
        s = take_slice(b, p0, p1)
        st, voff, vlen = verify_fn(s)
        n0 = int(b.proof_first[p0])
        byte0 = int(b.node_off[n0]) if int(b.proof_first[p1]) > n0 else 0
        voff = np.where(st == 0, voff + np.uint64(byte0), np.uint64(0)).astype(np.uint64)
    else:
        st, voff, vlen = np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    # gather: pad every rank's share to the largest slice (13 bytes per proof; not a data-path collective)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    m = max(cuts[r + 1] - cuts[r] for r in range(world))
    pack = np.zeros((m, 13), np.uint8)
    k = p1 - p0
    pack[:k, 0] = st
    pack[:k, 1:9] = voff.view(np.uint8).reshape(k, 8)
    pack[:k, 9:13] = vlen.view(np.uint8).reshape(k, 4)
    mine = torch.from_numpy(pack).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    status = np.zeros(b.n_proofs, np.uint8)
    value_off = np.zeros(b.n_proofs, np.uint64)
    value_len = np.zeros(b.n_proofs, np.uint32)
    for r in range(world):
        a = parts[r].cpu().numpy()
        k = cuts[r + 1] - cuts[r]
        status[cuts[r]:cuts[r + 1]] = a[:k, 0]
        value_off[cuts[r]:cuts[r + 1]] = np.ascontiguousarray(a[:k, 1:9]).view(np.uint64).reshape(k)
        value_len[cuts[r]:cuts[r + 1]] = np.ascontiguousarray(a[:k, 9:13]).view(np.uint32).reshape(k)
    return status, value_off, value_len
