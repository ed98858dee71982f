"""The streamed borsh entry (mptv_verify_borsh): blobs in, verdicts out, results reported as slices of the
caller's blobs.  Same verdicts and value bytes as flatten-then-verify and as the reference ELF's golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _concat(blobs):
    off = np.zeros(len(blobs) + 1, np.uint64)
    np.cumsum([len(x) for x in blobs], out=off[1:])
    return np.frombuffer(b"".join(blobs) + b"\0", np.uint8), off


class _Pinned:
    """a page-locked copy of a byte array (mptv_alloc_pinned): with such blobs mptv_verify_borsh runs in pull mode --
    the device gathers the node bytes straight from them"""

    def __init__(self, buf, lead=0):
        import ctypes
        import zk_state_proofs_b200 as z
        self.lib = z.load_library()
        self.ptr = self.lib.mptv_alloc_pinned(len(buf) + lead + 64)
        assert self.ptr
        whole = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(len(buf) + lead + 64,))
        self.arr = whole[lead:lead + len(buf)]   # `lead` shifts every blob to another alignment
        self.arr[:] = buf

    def __del__(self):
        try:
            self.lib.mptv_free_pinned(self.ptr)
        except Exception:
            pass


@pytest.mark.parametrize("pinned", [0, 1, 2])
@pytest.mark.parametrize("host_dedup", [1, 0])
@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 16, 32 << 20])
def test_golden_vectors_through_the_borsh_stream(verifier, golden, chunk_bytes, host_dedup, pinned):
    """pinned = 1, 2: the blobs are page-locked (at two alignments), so the entry runs in PULL mode -- staging carries
    the index arrays and a gather list, k_gather fetches the node and key bytes from the blobs over PCIe"""
    import zk_state_proofs_b200 as z
    vs = golden["vectors"]
    blobs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh() for v in vs]
    buf, off = _concat(blobs)
    keep = None
    if pinned:
        keep = _Pinned(buf, lead=(0, 7)[pinned - 1])
        buf = keep.arr
    verifier.set_option("borsh_chunk_bytes", chunk_bytes)
    verifier.set_option("host_dedup", host_dedup)
    verifier.set_option("pull_pinned", 1)   # off by default (measured slower end to end); page-locked blobs are pulled when on
    verifier.host_stats(reset=True)
    try:
        st, voff, vlen = verifier.verify_borsh(buf, off, threads=4)
    finally:
        verifier.set_option("borsh_chunk_bytes", 32 << 20)
        verifier.set_option("host_dedup", 1)
        verifier.set_option("pull_pinned", 0)
    hs = verifier.host_stats()
    assert (hs.pull_chunks == hs.chunks) if pinned else (hs.pull_chunks == 0)   # page-locked blobs: every chunk is pulled
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
    pin = _Pinned(buf, lead=3)
    for dd, src in ((1, buf), (0, buf), (1, pin.arr), (0, pin.arr)):
        verifier.set_option("borsh_chunk_bytes", 1 << 20)
        verifier.set_option("host_dedup", dd)
        verifier.set_option("pull_pinned", 1)
        verifier.host_stats(reset=True)
        try:
            st2, voff2, vlen2 = verifier.verify_borsh(src, off)
        finally:
            verifier.set_option("borsh_chunk_bytes", 32 << 20)
            verifier.set_option("host_dedup", 1)
            verifier.set_option("pull_pinned", 0)
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


@pytest.mark.parametrize("lead", [0, 5])
@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 16, 32 << 20])
def test_device_flatten_mode(verifier, golden, oracle, chunk_bytes, lead):
    """borsh_mode 1: the page-locked blobs cross PCIe as they are and kernels flatten them (borsh_kernels.cu): same
    verdicts, the same value slices inside each proof's own blob as the host-flatten mode, BAD_ROOT_LEN, empty proofs,
    and a malformed blob fails the call"""
    import zk_state_proofs_b200 as z
    from oracle.fuzzgen import corpus
    vs = golden["vectors"]
    blobs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh() for v in vs]
    cases = corpus(4242 + lead, oracle.keccak256, 30, 1500, 2000, 1200, 1500, 600)
    blobs += [z.MerkleProofInput(c["proof"], c["root"], c["key"]).to_borsh() for c in cases]
    v0 = next(v for v in vs if v["status"] == 0)
    blobs += [z.MerkleProofInput(v0["proof_b"], v0["root_b"][:31], v0["key_b"]).to_borsh(), z.MerkleProofInput([], v0["root_b"], b"").to_borsh(),
              z.MerkleProofInput([b""], b"", b"").to_borsh()]
    buf, off = _concat(blobs)
    want = verifier.verify_borsh(buf, off, threads=4)          # host flatten, pageable
    pin = _Pinned(buf, lead=lead)
    verifier.set_option("borsh_mode", 1)
    verifier.set_option("borsh_chunk_bytes", chunk_bytes)
    verifier.host_stats(reset=True)
    try:
        got = verifier.verify_borsh(pin.arr, off)
        hs = verifier.host_stats()
        assert hs.device_chunks == hs.chunks > 0 and hs.nodes == sum(len(z.MerkleProofInput.from_borsh(b).proof) for b in blobs[-3:]) + \
            sum(len(v["proof_b"]) for v in vs) + sum(len(c["proof"]) for c in cases)
        for x, y in zip(got, want):
            assert (x == y).all()
        assert got[0][-3:].tolist() == [6, 1, 6]
        # hybrid: the host pipeline takes chunks from the front, the device pipeline from the back, at the same time
        verifier.set_option("borsh_mode", 2)
        verifier.host_stats(reset=True)
        both = verifier.verify_borsh(pin.arr, off, threads=4)
        hs = verifier.host_stats()
        assert all((x == y).all() for x, y in zip(both, want)) and hs.nodes == verifier.host_stats().nodes
        assert 0 < hs.device_chunks < hs.chunks or chunk_bytes >= (32 << 20)
        verifier.set_option("borsh_mode", 1)
        # pageable blobs fall back to the host flatten
        verifier.host_stats(reset=True)
        again = verifier.verify_borsh(buf, off, threads=4)
        assert verifier.host_stats().device_chunks == 0 and all((x == y).all() for x, y in zip(again, want))
        # what borsh::from_slice rejects fails the whole call, wherever it sits; the context stays usable
        good = blobs[0]
        for bad in (good[:-1], good + b"\0", good[:3], b"", b"\xff\xff\xff\xff" + good[4:]):
            b2, o2 = _concat([good] * 40 + [bad] + [good] * 40)
            p2 = _Pinned(b2, lead=lead)
            with pytest.raises(z.MptvError):
                verifier.verify_borsh(p2.arr, o2)
        st, _, _ = verifier.verify_borsh(_Pinned(_concat([good])[0]).arr, _concat([good])[1])
        assert st.tolist() == [vs[0]["status"]]
    finally:
        verifier.set_option("borsh_mode", -1)
        verifier.set_option("borsh_chunk_bytes", 32 << 20)


@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 20])
def test_write_combining_staging_gives_identical_results(verifier, golden, oracle, chunk_bytes):
    """"wc_staging": the chunk's node bytes are staged in write-combining page-locked memory (a block of their own, the
    index arrays stay where the result mapping can read them): same verdicts, same value slices, both streams"""
    import zk_state_proofs_b200 as z
    from oracle.fuzzgen import corpus
    from tests.test_gpu_storage_borsh import _inputs
    vs = golden["vectors"]
    blobs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh() for v in vs]
    blobs += [z.MerkleProofInput(c["proof"], c["root"], c["key"]).to_borsh() for c in corpus(99, oracle.keccak256, 30, 1500, 2000, 1200, 1500, 600)]
    buf, off = _concat(blobs)
    sblobs = [i.to_borsh() for i in _inputs(oracle, 5, 300)]
    verifier.set_option("borsh_chunk_bytes", chunk_bytes)
    try:
        for dd in (1, 0):
            verifier.set_option("host_dedup", dd)
            verifier.set_option("wc_staging", 0)
            want = verifier.verify_borsh(buf, off, threads=4)
            swant = verifier.verify_storage_borsh(sblobs, threads=3)
            verifier.set_option("wc_staging", 1)
            got = verifier.verify_borsh(buf, off, threads=4)
            sgot = verifier.verify_storage_borsh(sblobs, threads=3)
            assert all((x == y).all() for x, y in zip(got, want)) and all((x == y).all() for x, y in zip(sgot, swant))
            assert len(set(want[0].tolist())) >= 6
    finally:
        verifier.set_option("wc_staging", 0)
        verifier.set_option("host_dedup", 1)
        verifier.set_option("borsh_chunk_bytes", 32 << 20)


def test_device_flatten_parser_accepts_exactly_what_the_mirror_accepts(verifier, golden):
    """borsh_mode 1 on untrusted bytes: good blobs with corrupted length words, truncations, trailing bytes and bit flips.
    The device's walk over the length prefixes (k_blob_count) accepts a blob iff MerkleProofInput.from_borsh does; an
    accepted call gives the host-flatten mode's results"""
    import random
    import struct
    import zk_state_proofs_b200 as z
    vs = [v for v in golden["vectors"] if v["status"] == 0][:200]
    good = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh() for v in vs]
    rng = random.Random(3)
    verifier.set_option("borsh_chunk_bytes", 1 << 14)
    n_acc = n_rej = 0
    try:
        for it in range(400):
            b = bytearray(good[rng.randrange(len(good))])
            k = rng.random()
            if k < 0.35:
                p = rng.randrange(0, len(b) - 4)
                struct.pack_into("<I", b, p, rng.choice([0, 1, 2, 3, 0xffffffff, 0x7fffffff, len(b), len(b) - p, rng.randrange(0, 1 << 12)]))
            elif k < 0.5:
                b = b[:rng.randrange(0, len(b))]
            elif k < 0.6:
                b = b + bytes(rng.randrange(1, 40))
            elif k < 0.8:
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            else:
                struct.pack_into("<I", b, 0, rng.choice([0, 1, 2, 0xffffffff, rng.randrange(0, 64)]))
            b = bytes(b)
            try:
                z.MerkleProofInput.from_borsh(b)
                ok = True
            except Exception:
                ok = False
            at = rng.randrange(0, 60)
            buf, off = _concat(good[:at] + [b] + good[at:60])
            pin = _Pinned(buf, lead=it % 7)
            verifier.set_option("borsh_mode", 1)
            try:
                got = verifier.verify_borsh(pin.arr, off)
                assert ok, (it, len(b))
                n_acc += 1
                verifier.set_option("borsh_mode", 0)
                want = verifier.verify_borsh(buf, off, threads=2)
                assert all((x == y).all() for x, y in zip(got, want))
            except z.MptvError:
                assert not ok, (it, len(b))
                n_rej += 1
        assert n_acc > 60 and n_rej > 60
    finally:
        verifier.set_option("borsh_mode", -1)
        verifier.set_option("borsh_chunk_bytes", 32 << 20)
