"""The N > 1 path on CPU: world-size-2 `gloo` processes shard a batch as independent proof slices
(zk_state_proofs_b200.sharding), each verifies its slice, and the gathered per-proof results equal
a single-process run over the whole batch.  The per-rank "device" here is the CPU oracle (there is
no GPU in this container); on a GPU box the same code path runs with a Verifier per rank
(tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(nested):
    from workload import gen
    if nested:
        state, tokens = gen.make_state_and_tokens(20_000, 3, 5_000, seed=3)
        return gen.nested_batch(state, tokens, 1_500, seed=11)
    trie = gen.SynthTrie(30_000, 2, kind=0)
    return gen.account_batch(trie, 5_000, seed=9, p_excl=0.1, p_mut=0.2)


def _oracle_fn():
    from oracle.pyoracle import Oracle
    o = Oracle()

    def f(s):
        d = dict(node_bytes=s.node_bytes, node_off=s.node_off, node_len=s.node_len, proof_first=s.proof_first,
                 roots=s.roots, key_bytes=s.key_bytes, key_off=s.key_off)
        if s.root_from_proof is not None:
            d["root_from_proof"] = s.root_from_proof
        st, voff, vlen, _, _ = o.verify_batch(d, nthreads=1)
        return st, voff, vlen
    return f


def _worker(rank, world, port, nested, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zk_state_proofs_b200.sharding import verify_sharded
    b = _batch(nested)
    st, voff, vlen = verify_sharded(b, _oracle_fn())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), st=st, voff=voff, vlen=vlen)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nested", [False, True])
def test_world2_gloo_matches_single_process(tmp_path, nested):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nested, str(tmp_path)), nprocs=world, join=True)
    b = _batch(nested)
    st, voff, vlen = _oracle_fn()(b)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert (z["st"] == st).all() and (z["voff"] == voff).all() and (z["vlen"] == vlen).all()
    assert len(set(st.tolist())) >= 3


def test_cuts_balance_bytes_and_keep_groups():
    from zk_state_proofs_b200.sharding import slice_cuts, take_slice
    b = _batch(True)
    for world in (1, 2, 3, 8):
        cuts = slice_cuts(b, world)
        assert cuts[0] == 0 and cuts[-1] == b.n_proofs and all(x <= y for x, y in zip(cuts, cuts[1:]))
        sizes = []
        for r in range(world):
            p0, p1 = cuts[r], cuts[r + 1]
            assert p0 == b.n_proofs or b.root_from_proof[p0] < 0   # a slice starts at an account proof
            s = take_slice(b, p0, p1)
            assert s.n_proofs == p1 - p0 and (s.node_off % 16 == 0).all()
            assert s.root_from_proof is None or ((s.root_from_proof >= -1).all() and (s.root_from_proof < s.n_proofs).all())
            sizes.append(int(s.node_len.astype(np.int64).sum()))
        assert sum(sizes) == int(b.node_len.astype(np.int64).sum())
        if world > 1:
            assert max(sizes) - min(sizes) < 0.02 * sum(sizes) + 40_000
    # empty batch
    import zk_state_proofs_b200 as z
    assert slice_cuts(z.flatten([]), 4) == [0, 0, 0, 0, 0]
