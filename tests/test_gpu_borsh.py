"""The streamed borsh entry (mptv_verify_borsh): blobs in, verdicts out, results reported as slices of the
caller's blobs.  Same verdicts and value bytes as flatten-then-verify and as the reference ELF's golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _concat(blobs):
    off = np.zeros(len(blobs) + 1, np.uint64)
    np.cumsum([len(x) for x in blobs], out=off[1:])
    return np.frombuffer(b"".join(blobs) + b"\0", np.uint8), off


@pytest.mark.parametrize("host_dedup", [1, 0])
@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 16, 32 << 20])
def test_golden_vectors_through_the_borsh_stream(verifier, golden, chunk_bytes, host_dedup):
    import zk_state_proofs_b200 as z
    vs = golden["vectors"]
    blobs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh() for v in vs]
    buf, off = _concat(blobs)
    verifier.set_option("borsh_chunk_bytes", chunk_bytes)
    verifier.set_option("host_dedup", host_dedup)
    try:
        st, voff, vlen = verifier.verify_borsh(buf, off, threads=4)
    finally:
        verifier.set_option("borsh_chunk_bytes", 32 << 20)
        verifier.set_option("host_dedup", 1)
    assert st.tolist() == [v["expect_status"] for v in vs]
    for v, s, o, l in zip(vs, st, voff, vlen):
        if s == 0:
            assert buf[int(o):int(o) + int(l)].tobytes() == v["value_b"], v["tag"]
        else:
            assert (int(o), int(l)) == (0, 0)
    # the list-of-bytes form of the binding gives the same answer
    st2, voff2, vlen2 = verifier.verify_borsh(blobs[:300])
    assert (st2 == st[:300]).all() and (voff2 == voff[:300]).all() and (vlen2 == vlen[:300]).all()


def test_fuzz_corpus_stream_equals_flatten_then_verify(verifier, oracle):
    """the aliased transfer (host_dedup, the default) and the plain one give what flatten-then-verify gives, and the
    value slices point into each proof's OWN blob"""
    import zk_state_proofs_b200 as z
    from oracle.fuzzgen import corpus
    cases = corpus(31337, oracle.keccak256, 80, 4000, 5000, 3000, 4000, 1500)
    blobs = [z.MerkleProofInput(c["proof"], c["root"], c["key"]).to_borsh() for c in cases]
    buf, off = _concat(blobs)
    b = z.flatten_borsh(blobs)
    st, voff, vlen = verifier.verify_batch(b)
    aliased = {}
    for dd in (1, 0):
        verifier.set_option("borsh_chunk_bytes", 1 << 20)
        verifier.set_option("host_dedup", dd)
        verifier.host_stats(reset=True)
        try:
            st2, voff2, vlen2 = verifier.verify_borsh(buf, off)
        finally:
            verifier.set_option("borsh_chunk_bytes", 32 << 20)
            verifier.set_option("host_dedup", 1)
        hs = verifier.host_stats()
        aliased[dd] = (hs.nodes_aliased, hs.h2d_bytes)
        assert hs.nodes == b.n_nodes and hs.d2h_bytes == 13 * len(blobs)
        assert (st == st2).all() and (vlen == vlen2).all()
        for i in np.nonzero(st == 0)[0]:
            assert buf[int(voff2[i]):int(voff2[i]) + int(vlen2[i])].tobytes() == b.value(int(voff[i]), int(vlen[i]))
            assert int(off[i]) <= int(voff2[i]) and int(voff2[i]) + int(vlen2[i]) <= int(off[i + 1])
    assert len(set(st.tolist())) >= 6
    assert aliased[0][0] == 0 and aliased[1][0] > 0 and aliased[1][1] < aliased[0][1]


def test_bad_root_length_empty_and_malformed_blobs(verifier, golden):
    import zk_state_proofs_b200 as z
    v = next(v for v in golden["vectors"] if v["status"] == 0)
    good = z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh()
    short_root = z.MerkleProofInput(v["proof_b"], v["root_b"][:31], v["key_b"]).to_borsh()
    empty = z.MerkleProofInput([], v["root_b"], b"").to_borsh()
    st, voff, vlen = verifier.verify_borsh([good, short_root, empty, good])
    assert st.tolist() == [0, 6, 1, 0]
    cat = good + short_root + empty + good
    assert cat[int(voff[3]):int(voff[3]) + int(vlen[3])] == v["value_b"]
    assert cat[int(voff[0]):int(voff[0]) + int(vlen[0])] == v["value_b"] and voff[3] != voff[0]
    st, _, _ = verifier.verify_borsh([])
    assert len(st) == 0
    for bad in (good[:-1], good + b"\0", good[:3], b""):
        with pytest.raises(z.MptvError):
            verifier.verify_borsh([good, bad])
    # and the context is usable afterwards
    st, _, _ = verifier.verify_borsh([good])
    assert st.tolist() == [0]
