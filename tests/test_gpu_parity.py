"""Parity tests proper: the CUDA path, called through the C ABI (libmptv.so), against
(a) golden vectors produced by the reference's own ELF and (b) the C restatement on seeded inputs.
Bit-exact: verdict class, value bytes, value offsets, digests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(z, vs):
    return [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]) for v in vs]


def test_keccak_digests_match_oracle(verifier, oracle):
    rng = np.random.default_rng(1)
    lens = list(range(0, 300)) + [135, 136, 137, 271, 272, 273, 407, 408, 409, 532, 543, 544, 545, 1000, 4096,
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
def test_golden_vectors_from_reference_elf(verifier, golden, lanes):
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


def test_seeded_corpus_matches_oracle_including_offsets(verifier, oracle):
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


def test_small_chunks_and_no_binning_give_identical_results(verifier, golden):
    import zk_state_proofs_b200 as z
    vs = golden["vectors"]
    b = z.flatten(_inputs(z, vs))
    ref = verifier.verify_batch(b)
    verifier.set_option("chunk_bytes", 1 << 16)
    a = verifier.verify_batch(b)
    verifier.set_option("binning", 0)
    c = verifier.verify_batch(b)
    verifier.set_option("binning", 1)
    verifier.set_option("chunk_bytes", 96 << 20)
    for x, y, w in zip(ref, a, c):
        assert (x == y).all() and (x == w).all()


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


def test_empty_and_degenerate_batches(verifier):
    import zk_state_proofs_b200 as z
    assert verifier.verify_merkle_proofs([]) == []
    r = verifier.verify_merkle_proofs([z.MerkleProofInput([], b"\x11" * 32, b"\x01"),
                                       z.MerkleProofInput([b""], b"\x11" * 32, b""),
                                       z.MerkleProofInput([b"\x80"], bytes.fromhex(
                                           "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"), b"\x01")])
    assert [x.status for x in r] == [1, 1, 4]
