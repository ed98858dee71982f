"""mptv_verify_storage_borsh: borsh(StorageProofInput) blobs in (crypto-ops/src/types.rs:11-19), the storage guest's
flow out (storage-circuit/src/main.rs:6-31), against the oracle's restatement of that flow proof by proof and against
the flatten-then-verify entry (mptv_verify_batch_hashed_keys)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(oracle, seed, n_groups=400):
    """StorageProofInputs built from generated tries: real account leaves carrying the storage tries' roots, raw slots
    as storage keys, and every way the zip / the decode / a proof can go wrong"""
    import random
    import zk_state_proofs_b200 as z
    from oracle.pytrie import Trie, rlp_list, rlp_str, rlp_uint
    rng = random.Random(seed)
    k256 = oracle.keccak256
    tokens = []
    for t in range(4):
        slots = [rng.randbytes(32) for _ in range(rng.choice([1, 7, 60, 300]))]
        tr = Trie({k256(s): rlp_uint(rng.randrange(1, 1 << rng.choice([7, 8, 64, 200]))) for s in slots}, k256)
        tokens.append((tr, slots))
    # the state trie: accounts whose storage_root is one of the token roots; some leaves are not Accounts
    kv, accounts = {}, []
    for a in range(300):
        addr = rng.randbytes(20)
        tok = rng.randrange(len(tokens))
        kind = rng.random()
        if kind < 0.9:
            val = rlp_list([rlp_uint(rng.randrange(0, 1 << 20)), rlp_uint(rng.randrange(0, 1 << 100)), rlp_str(tokens[tok][0].root),
                            rlp_str(rng.randbytes(32))])
        elif kind < 0.95:
            val = rlp_list([b"\x01", b"\x02", rlp_str(tokens[tok][0].root[:31]), rlp_str(rng.randbytes(32))])  # storage_root of 31 bytes
        else:
            val = rlp_str(rng.randbytes(40))  # a string, not a list
        kv[k256(addr)] = val
        accounts.append((addr, tok))
    state = Trie(kv, k256)
    root = state.root
    inputs = []
    for g in range(n_groups):
        addr, tok = accounts[rng.randrange(len(accounts))]
        tr, slots = tokens[tok]
        ak = k256(addr) if rng.random() > 0.04 else k256(rng.randbytes(20))  # absent account now and then
        aproof = state.proof(ak)
        ns = rng.choice([0, 1, 3, 3, 5])
        keys, proofs = [], []
        for _ in range(ns):
            s = slots[rng.randrange(len(slots))] if rng.random() > 0.15 else rng.randbytes(32)  # absent slots too
            if rng.random() < 0.05:
                s = s[:rng.randrange(0, 31)]  # a short raw key: still hashed to 32 bytes
            pr = tr.proof(k256(s))
            m = rng.random()
            if m < 0.05 and pr:
                pr = pr[:-1]
            elif m < 0.10 and pr:
                i = rng.randrange(len(pr))
                b = bytearray(pr[i]); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8); pr = pr[:i] + [bytes(b)] + pr[i + 1:]
            elif m < 0.15:
                pr = pr[::-1]  # order must not matter
            keys.append(s)
            proofs.append(pr)
        z_ = rng.random()
        if z_ < 0.1 and proofs:
            proofs = proofs + [tr.proof(k256(slots[0]))] * rng.randrange(1, 3)  # more proofs than keys: the zip drops them
        elif z_ < 0.2 and keys:
            keys = keys + [rng.randbytes(32)] * rng.randrange(1, 3)               # more keys than proofs
        rh = root if rng.random() > 0.03 else root[:rng.choice([0, 31])] + b"\0" * rng.choice([0, 2])  # try_into().unwrap() fails
        inputs.append(z.StorageProofInput(aproof, proofs, rh, rng.randbytes(rng.choice([0, 20, 32])), keys, ak))
    return inputs


def _guest_flow(oracle, inputs):
    """the storage guest restated with the oracle as verify_merkle_proof (tests/test_gpu_storage.py::_guest)"""
    import zk_state_proofs_b200 as z
    from tests.test_gpu_storage import _guest
    out = []
    for inp in inputs:
        r = 6 if len(inp.root_hash) != 32 else _guest(oracle, inp)  # try_into().unwrap() comes first (main.rs:11)
        out.append(z.VerifyPanic(r) if isinstance(r, int) else r)
    return out


@pytest.mark.parametrize("chunk_bytes", [1 << 12, 1 << 16, 32 << 20])
@pytest.mark.parametrize("dedup", [1, 0])
def test_storage_borsh_stream_matches_the_guest_flow(verifier, oracle, chunk_bytes, dedup):
    import zk_state_proofs_b200 as z
    inputs = _inputs(oracle, 77 + dedup)
    want = _guest_flow(oracle, inputs)
    assert [w if isinstance(w, list) else w.status for w in want] == \
        [w if isinstance(w, list) else w.status for w in verifier.verify_storage_proof_inputs(inputs)]  # the CSR mirror agrees
    verifier.set_option("borsh_chunk_bytes", chunk_bytes)
    verifier.set_option("host_dedup", dedup)
    try:
        blobs = [i.to_borsh() for i in inputs]
        pf, ist, st, voff, vlen = verifier.verify_storage_borsh(blobs, threads=3)
        buf = np.frombuffer(b"".join(blobs), np.uint8)
        n_used = [1 + min(len(i.storage_proofs), len(i.storage_keys)) for i in inputs]
        assert (np.diff(pf.astype(np.int64)) == np.array(n_used)).all() and len(st) == sum(n_used)
        seen = set()
        for i, (inp, w) in enumerate(zip(inputs, want)):
            a, e = int(pf[i]), int(pf[i + 1])
            if isinstance(w, z.VerifyPanic):
                assert int(ist[i]) == w.status, (i, int(ist[i]), w.status, st[a:e].tolist())
                seen.add(int(ist[i]))
            else:
                assert ist[i] == 0 and (st[a:e] == 0).all()
                got = [buf[int(voff[q]):int(voff[q]) + int(vlen[q])].tobytes() for q in range(a + 1, e)]
                assert got == w
                # every value lies inside the input's own blob
                lo, hi = sum(len(b) for b in blobs[:i]), sum(len(b) for b in blobs[:i + 1])
                assert all(lo <= int(voff[q]) and int(voff[q]) + int(vlen[q]) <= hi for q in range(a, e))
                seen.add(0)
        assert {0, 1, 3, 4, 6, 7} <= seen | {int(s) for s in st}
        # the mirror that goes through the wire format gives what the CSR mirror gives
        via = verifier.verify_storage_proof_inputs_borsh(inputs[:120])
        for x, y, inp in zip(via, want[:120], inputs[:120]):
            if isinstance(y, z.VerifyPanic):
                assert isinstance(x, z.VerifyPanic)
            else:
                assert x == y
    finally:
        verifier.set_option("borsh_chunk_bytes", 32 << 20)
        verifier.set_option("host_dedup", 1)


def test_storage_borsh_account_leaf_must_decode_even_without_storage_proofs(verifier, oracle):
    import zk_state_proofs_b200 as z
    from oracle.pytrie import Trie, rlp_list, rlp_str, rlp_uint
    k256 = oracle.keccak256
    good = rlp_list([rlp_uint(5), rlp_uint(256), rlp_str(b"\x11" * 32), rlp_str(b"\x22" * 32)])
    vals = {b"a" * 20: good,
            b"b" * 20: rlp_list([rlp_uint(5), rlp_str(b"\x00\x01"), rlp_str(b"\x11" * 32), rlp_str(b"\x22" * 32)]),  # balance with a leading zero
            b"c" * 20: good,
            b"d" * 20: rlp_list([rlp_uint(5), rlp_uint(1), rlp_str(b"\x11" * 32), rlp_str(b"\x22" * 32), b"\x80"])}  # five items
    tr = Trie({k256(a): v for a, v in vals.items()}, k256)
    ins = [z.StorageProofInput(tr.proof(k256(a)), [], tr.root, b"", [], k256(a)) for a in vals]
    pf, ist, st, voff, vlen = verifier.verify_storage_borsh([i.to_borsh() for i in ins])
    assert (st == 0).all() and ist.tolist() == [0, 7, 0, 7] and pf.tolist() == [0, 1, 2, 3, 4]


def test_storage_borsh_malformed_blobs_fail_the_call(verifier, oracle):
    import zk_state_proofs_b200 as z
    inputs = _inputs(oracle, 5, n_groups=60)
    blobs = [i.to_borsh() for i in inputs]
    good = next(b for b, i in zip(blobs, inputs) if len(i.storage_proofs) >= 2)
    for bad in (good[:-1], good + b"\0", good[:3], b"", b"\xff\xff\xff\xff" + good[4:], good[:-33]):
        with pytest.raises(z.MptvError):
            verifier.verify_storage_borsh(blobs[:30] + [bad] + blobs[30:])
    pf, ist, st, _, _ = verifier.verify_storage_borsh(blobs)  # the context stays usable
    assert len(st) == int(pf[-1])
    assert verifier.verify_storage_borsh([])[0].tolist() == [0]


def test_storage_borsh_stream_at_scale_equals_the_csr_batch(verifier):
    """the nested synthetic workload (config 3 shape: 80 / 10 / 10, all mutators) written out as StorageProofInput blobs
    and streamed back in gives, proof for proof, what the flattened batch gives through mptv_verify_batch"""
    from workload import gen
    state, tokens = gen.make_state_and_tokens(200_000, 3, 20_000, seed=3)
    nb = gen.nested_batch(state, tokens, 20_000, seed=9, raw_keys=True)
    blobs, off, gf = gen.batch_to_storage_borsh(nb)
    st, voff, vlen = verifier.verify_batch(nb)
    assert len(set(st.tolist())) >= 5
    for chunk in (1 << 20, 32 << 20):
        verifier.set_option("borsh_chunk_bytes", chunk)
        pf, ist, bst, bvoff, bvlen = verifier.verify_storage_borsh(blobs, off)
        assert (pf == gf).all() and (bst == st).all() and (bvlen == vlen).all()
        for i in np.nonzero(bst == 0)[0][::13]:
            assert blobs[int(bvoff[i]):int(bvoff[i]) + int(bvlen[i])].tobytes() == nb.value(int(voff[i]), int(vlen[i]))
        for g in range(0, len(gf) - 1, 7):
            assert int(ist[g]) == next((int(s) for s in st[int(gf[g]):int(gf[g + 1])] if s), 0)
    # the host half on its own: mptv_flatten_storage_borsh -> mptv_verify_batch_hashed_keys gives the same verdicts and values
    import zk_state_proofs_b200 as z
    fb, hk, fpf, info = z.flatten_storage_borsh(blobs, off)
    assert (fpf == gf).all() and int(hk.sum()) == nb.n_proofs - (len(gf) - 1)
    fst, fvoff, fvlen = verifier.verify_batch_hashed_keys(fb, hk)
    assert (fst == st).all() and (fvlen == vlen).all()
    for i in np.nonzero(fst == 0)[0][::29]:
        assert fb.value(int(fvoff[i]), int(fvlen[i])) == nb.value(int(voff[i]), int(vlen[i]))
    # a result capacity that is too small (the stream stops at the chunk that does not fit and reports the real count),
    # exactly right, and too large
    verifier.set_option("borsh_chunk_bytes", 1 << 20)
    for cap in (5, nb.n_proofs // 2, nb.n_proofs, nb.n_proofs + 1000):
        pf, ist, bst, bvoff, bvlen = verifier.verify_storage_borsh(blobs, off, n_proofs=cap)
        assert (pf == gf).all() and len(bst) == nb.n_proofs and (bst == st).all() and (bvlen == vlen).all()
    verifier.set_option("borsh_chunk_bytes", 32 << 20)
    hs = verifier.host_stats()
    assert hs.nodes_aliased > 0 and hs.index_us > 0
