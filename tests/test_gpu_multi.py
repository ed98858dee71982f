"""N > 1 on real GPUs (skipped on a 1-GPU box): (a) one context over two devices -- the C ABI's own
slice-per-device path -- and (b) one process per GPU under torchrun with the NCCL gather of
zk_state_proofs_b200.sharding.verify_sharded, both against the single-GPU result."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _batches():
    from workload import gen
    state, tokens = gen.make_state_and_tokens(200_000, 3, 20_000, seed=3)
    return gen.nested_batch(state, tokens, 10_000, seed=21), gen.block_tries(40, 300, "both", seed=4)


def test_one_context_two_devices_matches_one_device(verifier):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import zk_state_proofs_b200 as z
    b, kv = _batches()
    one = verifier.verify_batch(b)
    two = z.Verifier([0, 1])
    assert two.device_count == 2
    two.set_option("chunk_bytes", 4 << 20)
    got = two.verify_batch(b)
    for x, y in zip(one, got):
        assert (x == y).all()
    assert (two.trie_roots(kv) == verifier.trie_roots(kv)).all()
    # rebuild + get_proof: trie-ordered targets are cut over the two devices with their tries, unordered ones go to device 0
    import zk_state_proofs_b200 as z2
    keys = [z2.rlp_index(i) for i in (0, 1, 15, 127, 128, 299, 300)]
    ordered = [(t, k) for t in range(kv.n_tries) for k in keys]
    shuffled = ordered[::-1]
    for targets in (ordered, shuffled):
        r1, b1 = verifier.trie_proofs(kv, targets)
        r2, b2 = two.trie_proofs(kv, targets)
        assert (r1 == r2).all() and (b1.proof_first == b2.proof_first).all() and (b1.node_len == b2.node_len).all()
        assert (b1.node_off == b2.node_off).all() and (b1.node_bytes == b2.node_bytes).all()
        s1 = verifier.verify_batch(b1)
        s2 = two.verify_batch(b2)
        assert all((x == y).all() for x, y in zip(s1, s2)) and set(s1[0].tolist()) == {0, 4}
    # the streamed borsh entry over two devices (blobs cut by bytes, one host thread pool per device)
    from workload import gen
    acc = gen.account_batch(gen.SynthTrie(300_000, 2, kind=0), 30_000, seed=5)
    blobs, off = gen.batch_to_borsh(acc)
    two.set_option("borsh_chunk_bytes", 2 << 20)
    s1, o1, l1 = verifier.verify_borsh(blobs, off)
    s2, o2, l2 = two.verify_borsh(blobs, off)
    assert (s1 == s2).all() and (o1 == o2).all() and (l1 == l2).all() and (s1 == 0).all()
    ref = verifier.verify_batch(acc)
    for i in range(0, acc.n_proofs, 97):
        assert blobs[int(o2[i]):int(o2[i]) + int(l2[i])].tobytes() == acc.value(int(ref[1][i]), int(ref[2][i]))
    # the same blobs page-locked: a context with several devices flattens them on the devices by itself (borsh_mode -1)
    pblobs, poff = gen.batch_to_borsh(acc, pinned=True)
    two.host_stats(reset=True)
    s3, o3, l3 = two.verify_borsh(pblobs, poff)
    hs = two.host_stats()
    assert hs.device_chunks == hs.chunks > 1 and (s3 == s1).all() and (o3 == o1).all() and (l3 == l1).all()
    two.set_option("borsh_mode", 0)
    two.host_stats(reset=True)
    s4, o4, l4 = two.verify_borsh(pblobs, poff)
    assert two.host_stats().device_chunks == 0 and (s4 == s1).all() and (o4 == o1).all() and (l4 == l1).all()
    two.set_option("borsh_mode", -1)
    # digest_keccak over an arena, cut into index ranges over the two devices (mptv_keccak256_batch), and the
    # hashed-keys entry on two devices
    d1 = verifier.keccak256_batch(acc.node_bytes, acc.node_off, acc.node_len)
    d2 = two.keccak256_batch(acc.node_bytes, acc.node_off, acc.node_len)
    assert (d1 == d2).all() and len(d2) == acc.n_nodes
    flags = np.zeros(acc.n_proofs, np.uint8)
    h1 = verifier.verify_batch_hashed_keys(acc, flags)
    h2 = two.verify_batch_hashed_keys(acc, flags)
    assert all((x == y).all() for x, y in zip(h1, h2)) and (h2[0] == 0).all()
    # the storage guest's wire format over two devices (inputs cut by bytes at input boundaries)
    state, tokens = gen.make_state_and_tokens(100_000, 3, 10_000, seed=3)
    nb = gen.nested_batch(state, tokens, 8_000, seed=9, raw_keys=True)
    sblobs, soff, gf = gen.batch_to_storage_borsh(nb)
    two.set_option("borsh_chunk_bytes", 1 << 20)
    r1 = verifier.verify_storage_borsh(sblobs, soff)
    r2 = two.verify_storage_borsh(sblobs, soff)
    assert all((x == y).all() for x, y in zip(r1, r2)) and (r2[0] == gf).all() and len(set(r2[2].tolist())) >= 5
    two.close()


def test_torchrun_two_ranks_nccl_matches_one_device(verifier, tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    b, _ = _batches()
    st, voff, vlen = verifier.verify_batch(b)
    out = str(tmp_path / "sharded.npz")
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29517",
                           os.path.join(ROOT, "tests", "_sharded_nccl.py"), out], env=env, timeout=600)
    z = np.load(out)
    assert (z["st"] == st).all() and (z["voff"] == voff).all() and (z["vlen"] == vlen).all()
