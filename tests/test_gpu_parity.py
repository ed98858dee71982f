"""Parity tests proper: the CUDA path, called through the C ABI (libmptv.so), against
(a) golden vectors produced by the reference's own ELF and (b) the C restatement on seeded inputs.
Bit-exact: verdict class, value bytes, value offsets, digests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[(1, 0), (0, 0), (1, 1)], ids=["fast_walk", "cooperative_walk_only", "fast_walk+dedup_nodes"])
def walk_mode(request, verifier):
    """walk configurations: K2f (thread per proof) + K2b on the deferred rest, K2b on everything, and the
    optional node de-duplication in front of K1 (every distinct node hashed once, digests shared)"""
    verifier.set_option("fast_walk", request.param[0])
    verifier.set_option("dedup_nodes", request.param[1])
    yield request.param
    verifier.set_option("fast_walk", 1)
    verifier.set_option("dedup_nodes", 0)


def _inputs(z, vs):
    return [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]) for v in vs]


def test_keccak_digests_match_oracle(verifier, oracle):
    rng = np.random.default_rng(1)
    lens = list(range(0, 1100)) + [4096, 4487, 4488, 4489,
                                  8191, 30000] + list(rng.integers(0, 2000, 500))
    lens = np.array(lens, np.uint32)
    rng.shuffle(lens)
    padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
    off = np.zeros(len(lens), np.uint64)
    np.cumsum(padded[:-1], out=off[1:])
    nb = rng.integers(0, 256, int(padded.sum()) + 16, dtype=np.uint8)
    got = verifier.keccak256_batch(nb, off, lens)
    want = oracle.keccak256_batch(nb, off, lens)
    assert (got == want).all(), np.nonzero((got != want).any(axis=1))[0][:10]
    # empty-string and empty-trie KATs
    assert verifier.digest_keccak(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert verifier.digest_keccak(b"\x80").hex() == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"


@pytest.mark.parametrize("lanes", [0, 8, 16, 32])
def test_golden_vectors_from_reference_elf(verifier, golden, lanes, walk_mode):
    import zk_state_proofs_b200 as z
    verifier.set_option("lanes_per_proof", lanes)
    vs = golden["vectors"]
    res = verifier.verify_merkle_proofs(_inputs(z, vs))
    bad = []
    for v, r in zip(vs, res):
        got_status = r.status if isinstance(r, z.VerifyPanic) else 0
        got_value = None if isinstance(r, z.VerifyPanic) else r
        if got_status != v["status"] or got_value != v["value_b"]:
            bad.append((v["tag"], got_status, v["status"]))
    verifier.set_option("lanes_per_proof", 0)
    assert not bad, (len(bad), bad[:10])


def test_seeded_corpus_matches_oracle_including_offsets(verifier, oracle, walk_mode):
    import zk_state_proofs_b200 as z
    from oracle.fuzzgen import corpus
    cases = corpus(2024, oracle.keccak256, 60, 1500, 3000)
    b = z.flatten([z.MerkleProofInput(c["proof"], c["root"], c["key"]) for c in cases])
    st, voff, vlen = verifier.verify_batch(b)
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off)
    ost, ovoff, ovlen, _, _ = oracle.verify_batch(d, nthreads=4)
    assert (st == ost).all(), [(cases[i]["tag"], st[i], ost[i]) for i in np.nonzero(st != ost)[0][:10]]
    assert (vlen == ovlen).all()
    assert (voff == ovoff).all()


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_long_proofs_beyond_the_lane_group(verifier, oracle, lanes, walk_mode):
    """Proofs with far more nodes than lanes in a group (the reference takes ANY multiset of nodes): the
    path's own nodes scattered among up to 100 junk strings, foreign nodes of other proofs and duplicates."""
    import random
    import zk_state_proofs_b200 as z
    from oracle.fuzzgen import corpus
    rng = random.Random(77)
    base = corpus(4242, oracle.keccak256, 40, 400, 300)
    pool = [n for c in base for n in c["proof"] if n]
    cases = []
    for c in base:
        proof = list(c["proof"])
        for _ in range(rng.choice([0, 3, 7, 9, 15, 17, 31, 33, 60, 100])):
            r = rng.random()
            extra = rng.randbytes(rng.randint(1, 80)) if r < 0.4 else (rng.choice(pool) if r < 0.8 or not proof
                                                                       else rng.choice(proof))
            proof.insert(rng.randrange(len(proof) + 1), extra)
        if rng.random() < 0.5:
            rng.shuffle(proof)
        cases.append(dict(root=c["root"], proof=proof, key=c["key"], tag=c["tag"]))
    b = z.flatten([z.MerkleProofInput(c["proof"], c["root"], c["key"]) for c in cases])
    verifier.set_option("lanes_per_proof", lanes)
    try:
        st, voff, vlen = verifier.verify_batch(b)
    finally:
        verifier.set_option("lanes_per_proof", 0)
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off)
    ost, ovoff, ovlen, _, _ = oracle.verify_batch(d, nthreads=4)
    assert (st == ost).all(), [(cases[i]["tag"], len(cases[i]["proof"]), st[i], ost[i]) for i in np.nonzero(st != ost)[0][:10]]
    # duplicates of the node that holds the value: both sides report the first copy in proof order
    assert (vlen == ovlen).all() and (voff == ovoff).all()
    assert (np.array([len(c["proof"]) for c in cases]) > 32).sum() > 50


def test_small_chunks_no_binning_unfused_give_identical_results(verifier, golden):
    import zk_state_proofs_b200 as z
    vs = golden["vectors"]
    b = z.flatten(_inputs(z, vs))
    ref = verifier.verify_batch(b)
    verifier.set_option("chunk_bytes", 1 << 16)
    a = verifier.verify_batch(b)
    verifier.set_option("binning", 0)
    c = verifier.verify_batch(b)
    verifier.set_option("binning", 1)
    verifier.set_option("fused_classify", 0)  # every node decided by k_parse_nodes instead of K1's fast path
    d = verifier.verify_batch(b)
    verifier.set_option("fused_classify", 1)
    verifier.set_option("chunk_bytes", 96 << 20)
    for x, y, w, v in zip(ref, a, c, d):
        assert (x == y).all() and (x == w).all() and (x == v).all()


def test_synthetic_state_proofs_mixed_match_oracle(verifier, oracle, walk_mode):
    """config-2/3 shaped batch (inclusion, exclusion, all 7 mutators) at a size the oracle does in seconds."""
    from workload import gen
    trie = gen.SynthTrie(300_000, 2, kind=0)
    b = gen.account_batch(trie, 60_000, seed=5, p_excl=0.10, p_mut=0.20)
    st, voff, vlen = verifier.verify_batch(b)
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off)
    ost, ovoff, ovlen, pa, _ = oracle.verify_batch(d, nthreads=8)
    assert pa == b.n_perm()
    assert (st == ost).all() and (voff == ovoff).all() and (vlen == ovlen).all()
    counts = np.bincount(st, minlength=8)
    assert counts[0] > 40_000 and counts[1] > 0 and counts[3] > 0 and counts[4] > 0
    # the returned value of an accepted inclusion proof is the account RLP the generator inserted
    k, v = trie.entry(0)
    one = verifier.verify_merkle_proof(trie.root.tobytes(), _proof_of(b, 0), b.key_bytes[:32].tobytes()) if st[0] == 0 else None
    assert one is None or len(one) >= 70


def _proof_of(b, p):
    return [b.node_bytes[int(b.node_off[i]):int(b.node_off[i]) + int(b.node_len[i])].tobytes()
            for i in range(int(b.proof_first[p]), int(b.proof_first[p + 1]))]


def test_nested_account_storage_groups_match_oracle(verifier, oracle, walk_mode):
    """config 3: storage proofs take their root from the verified account leaf (root_from_proof)."""
    from workload import gen
    state, tokens = gen.make_state_and_tokens(100_000, 4, 50_000, seed=3)
    b = gen.nested_batch(state, tokens, 15_000, seed=3)
    st, voff, vlen = verifier.verify_batch(b)
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off, root_from_proof=b.root_from_proof)
    ost, ovoff, ovlen, _, _ = oracle.verify_batch(d, nthreads=8)
    assert (st == ost).all() and (voff == ovoff).all() and (vlen == ovlen).all()
    counts = np.bincount(st, minlength=8)
    assert counts[0] > 30_000 and counts[7] > 0 and counts[4] > 0
    # chunked pipeline must keep groups together
    verifier.set_option("chunk_bytes", 1 << 20)
    st2, voff2, vlen2 = verifier.verify_batch(b)
    verifier.set_option("chunk_bytes", 96 << 20)
    assert (st2 == st).all() and (voff2 == voff).all() and (vlen2 == vlen).all()


def test_single_proof_api_and_panics(verifier, golden):
    import zk_state_proofs_b200 as z
    v = next(v for v in golden["vectors"] if v["tag"] == "config1/tx15")
    assert verifier.verify_merkle_proof(v["root_b"], v["proof_b"], v["key_b"]) == v["value_b"]
    w = next(v for v in golden["vectors"] if v["tag"] == "config1/tx200")
    with pytest.raises(z.VerifyPanic) as e:
        verifier.verify_merkle_proof(w["root_b"], w["proof_b"], w["key_b"])
    assert e.value.status == 4
    with pytest.raises(z.VerifyPanic) as e:
        verifier.verify_merkle_proof(b"\x00" * 31, v["proof_b"], v["key_b"])
    assert e.value.status == 6


def test_dedup_counts_and_padding_is_not_compared(verifier, oracle):
    """dedup_nodes: the unique-node count is exact (nodes equal byte for byte are merged, padding garbage after a
    node never matters, near-duplicates are kept apart) and the results do not change"""
    import zk_state_proofs_b200 as z
    from workload import gen
    trie = gen.SynthTrie(200_000, 2, kind=0)
    b = gen.account_batch(trie, 40_000, seed=5, p_excl=0.05, p_mut=0.10)
    want = verifier.verify_batch(b)
    # scribble over the padding bytes between nodes: duplicates must still be found, results unchanged
    nb = b.node_bytes.copy()
    rng = np.random.default_rng(3)
    ends = (b.node_off + b.node_len.astype(np.uint64)).astype(np.int64)
    pad = (-b.node_len.astype(np.int64)) % 16
    for k in range(1, 16):
        sel = ends[pad >= k] + (k - 1)
        nb[sel] = rng.integers(0, 256, len(sel), dtype=np.uint8)
    b2 = z.Batch(nb, b.node_off, b.node_len, b.proof_first, b.roots, b.key_bytes, b.key_off, None, None)
    verifier.set_option("dedup_nodes", 1)
    try:
        got = verifier.verify_batch(b2)
        # exact number of distinct nodes, from the host
        seen = set()
        for o, n in zip(b.node_off.tolist(), b.node_len.tolist()):
            seen.add(b.node_bytes[o:o + n].tobytes())
        import torch
        dev = torch.device("cuda", 0)
        t = {k: torch.from_numpy(getattr(b2, k).view(np.uint8) if getattr(b2, k).dtype != np.uint8 else getattr(b2, k)).to(dev)
             for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_bytes", "key_off"]}
        d_st = torch.zeros(b2.n_proofs, dtype=torch.uint8, device=dev)
        d_vo = torch.zeros(b2.n_proofs, dtype=torch.int64, device=dev)
        d_vl = torch.zeros(b2.n_proofs, dtype=torch.int32, device=dev)
        verifier.verify_batch_device(0, {k: v.data_ptr() for k, v in t.items()}, b2.n_nodes, b2.n_proofs,
                                     dict(status=d_st.data_ptr(), value_off=d_vo.data_ptr(), value_len=d_vl.data_ptr()),
                                     node_bytes_len=len(b2.node_bytes))
        tm = verifier.last_timings(0)
        assert tm.n_unique_nodes == len(seen) < b.n_nodes
        assert (d_st.cpu().numpy() == want[0]).all()
    finally:
        verifier.set_option("dedup_nodes", 0)
    for x, y in zip(want, got):
        assert (x == y).all()


def test_empty_and_degenerate_batches(verifier):
    import zk_state_proofs_b200 as z
    assert verifier.verify_merkle_proofs([]) == []
    r = verifier.verify_merkle_proofs([z.MerkleProofInput([], b"\x11" * 32, b"\x01"),
                                       z.MerkleProofInput([b""], b"\x11" * 32, b""),
                                       z.MerkleProofInput([b"\x80"], bytes.fromhex(
                                           "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"), b"\x01")])
    assert [x.status for x in r] == [1, 1, 4]
