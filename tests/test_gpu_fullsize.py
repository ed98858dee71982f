"""BASELINE.json's full config-2 size (1 M account proofs against a 10 M-account state trie) through
size-independent properties, plus the oracle on a random sample:
  * every inclusion proof is accepted and returns an Account RLP
  * the proof is a SET: reversing the node order of every proof changes nothing (R6)
  * flipping one bit of every proof's last node rejects every proof as InvalidProof (R10)
  * dropping the root node rejects every proof as InvalidStateRoot (R2)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from workload import gen
    trie = gen.SynthTrie(10_000_000, 2, kind=0)
    b = gen.account_batch(trie, 1_000_000, seed=2)
    trie.close()
    return b


def _values(b, voff, vlen, idx):
    return [b.node_bytes[int(voff[i]):int(voff[i]) + int(vlen[i])].tobytes() for i in idx]


def test_full_size_inclusion_and_sample_against_oracle(verifier, oracle, big):
    st, voff, vlen = verifier.verify_batch(big)
    assert (st == 0).all()
    assert big.n_perm() > 23_000_000
    rng = np.random.default_rng(0)
    idx = rng.choice(big.n_proofs, 20_000, replace=False)
    for i in idx[:2000]:
        assert oracle.account_storage_root(big.value(int(voff[i]), int(vlen[i]))) is not None
    # oracle on the sampled proofs (re-assembled as a small batch)
    import zk_state_proofs_b200 as z
    from zk_state_proofs_b200.sharding import take_slice
    for i in idx[:3000]:
        s = take_slice(big, int(i), int(i) + 1)
        o = oracle.verify(s.roots.tobytes(), [s.node_bytes[int(a):int(a) + int(n)].tobytes() for a, n in zip(s.node_off, s.node_len)],
                          s.key_bytes[:int(s.key_off[1])].tobytes())
        assert o[0] == 0 and o[1] == big.value(int(voff[i]), int(vlen[i]))


def test_full_size_order_independence_and_tamper(verifier, big):
    import zk_state_proofs_b200 as z
    st, voff, vlen = verifier.verify_batch(big)
    n = big.n_nodes
    # reverse the node order inside every proof: node k of proof p <- node (first + last - k)
    first = np.repeat(big.proof_first[:-1].astype(np.int64), np.diff(big.proof_first.astype(np.int64)))
    last = np.repeat(big.proof_first[1:].astype(np.int64) - 1, np.diff(big.proof_first.astype(np.int64)))
    perm = first + last - np.arange(n, dtype=np.int64)
    rev = z.Batch(big.node_bytes, big.node_off[perm].copy(), big.node_len[perm].copy(), big.proof_first, big.roots,
                  big.key_bytes, big.key_off, None, None)
    # the host entry requires nodes in arena order, so hand the permuted index arrays to the device entry
    import torch
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(getattr(rev, k).view(np.uint8) if getattr(rev, k).dtype != np.uint8 else getattr(rev, k)).to(dev)
         for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_bytes", "key_off"]}
    d_st = torch.zeros(rev.n_proofs, dtype=torch.uint8, device=dev)
    d_vo = torch.zeros(rev.n_proofs, dtype=torch.int64, device=dev)
    d_vl = torch.zeros(rev.n_proofs, dtype=torch.int32, device=dev)
    verifier.verify_batch_device(0, {k: v.data_ptr() for k, v in t.items()}, rev.n_nodes, rev.n_proofs,
                                 dict(status=d_st.data_ptr(), value_off=d_vo.data_ptr(), value_len=d_vl.data_ptr()),
                                 node_bytes_len=len(rev.node_bytes))
    torch.cuda.synchronize()
    assert (d_st.cpu().numpy() == st).all()
    assert (d_vo.cpu().numpy().view(np.uint64) == voff).all() and (d_vl.cpu().numpy().view(np.uint32) == vlen).all()
    # flip one bit in the last byte of every proof's last node (the leaf) -> InvalidProof everywhere
    bad = big.node_bytes.copy()
    lastn = big.proof_first[1:].astype(np.int64) - 1
    pos = big.node_off[lastn].astype(np.int64) + big.node_len[lastn].astype(np.int64) - 1
    bad[pos] ^= 1
    st2, _, vl2 = verifier.verify_batch(z.Batch(bad, big.node_off, big.node_len, big.proof_first, big.roots, big.key_bytes,
                                                big.key_off, None, None))
    assert (st2 == 3).all() and (vl2 == 0).all()
    # drop the root node of every proof -> InvalidStateRoot everywhere
    keep = np.ones(n, bool)
    keep[big.proof_first[:-1].astype(np.int64)] = False
    pf = np.zeros(big.n_proofs + 1, np.uint32)
    np.cumsum(np.diff(big.proof_first.astype(np.int64)) - 1, out=pf[1:])
    st3, _, _ = verifier.verify_batch(z.Batch(big.node_bytes, big.node_off[keep].copy(), big.node_len[keep].copy(), pf,
                                              big.roots, big.key_bytes, big.key_off, None, None))
    assert (st3 == 1).all()
