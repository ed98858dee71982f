"""The latency path (csrc/single_kernels.cu): a call small enough for one CTA -- every single
crypto_ops::verify_merkle_proof call (/root/reference/crypto-ops/src/lib.rs:8-23; call shape
/root/reference/trie-utils/tests/transaction.rs:18-22, tests/storage.rs:53-79) -- is ONE kernel launch with inputs and
results in mapped page-locked memory.  Same rule set as the batch kernels, so: same verdicts, values and offsets on
the whole golden set run one proof per call, on small dependent groups, and at the size limits of the path."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _one(ver, z, v):
    r = ver.verify_merkle_proofs([z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"])])[0]
    return (r.status, None) if isinstance(r, z.VerifyPanic) else (0, r)


def test_every_golden_vector_one_proof_per_call(verifier, golden):
    """the reference's verdict and value for each vector, through ONE launch per call (launch count checked)"""
    import zk_state_proofs_b200 as z
    bad, small_calls = [], 0
    for v in golden["vectors"]:
        verifier.host_stats(reset=True)
        got = _one(verifier, z, v)
        hs = verifier.host_stats()
        fits = len(v["proof_b"]) <= 128 and sum((len(n) + 15) // 16 * 16 + 12 for n in v["proof_b"]) + len(v["key_b"]) + 200 < (60 << 10)
        if fits and len(v["root_b"]) == 32:
            small_calls += 1
            assert hs.launches == 1, (v["tag"], hs.launches)
        if got != (v["status"], v["value_b"]):
            bad.append((v["tag"], got[0], v["status"]))
    assert not bad, (len(bad), bad[:10])
    assert small_calls > 2500


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_latency_path_equals_batch_pipeline(verifier, golden, lanes):
    """identical (status, value_off, value_len) with the path on and off, in batches of 1 ... 32 proofs"""
    import zk_state_proofs_b200 as z
    rng = random.Random(lanes)
    vs = golden["vectors"]
    verifier.set_option("lanes_per_proof", lanes)
    try:
        for trial in range(120):
            k = rng.choice([1, 1, 2, 3, 5, 16, 31, 32])
            pick = [vs[rng.randrange(len(vs))] for _ in range(k)]
            b = z.flatten([z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]) for v in pick])
            verifier.set_option("latency_path", 1)
            verifier.host_stats(reset=True)
            a = verifier.verify_batch(b)
            launches = verifier.host_stats().launches
            verifier.set_option("latency_path", 0)
            c = verifier.verify_batch(b)
            assert all((x == y).all() for x, y in zip(a, c)), [v["tag"] for v in pick]
            if b.n_nodes <= 128 and len(b.node_bytes) < (40 << 10):
                assert launches in (1, 2)   # 2: a trailing proof without nodes starts a chunk of its own
    finally:
        verifier.set_option("latency_path", 1)
        verifier.set_option("lanes_per_proof", 0)


def test_limits_of_the_path(verifier, golden, oracle):
    """129 nodes, 33 proofs or a pack above 64 KiB fall back to the batch pipeline; 128 / 32 stay on the one launch"""
    import zk_state_proofs_b200 as z
    v = next(v for v in golden["vectors"] if v["tag"] == "config1/tx15")
    junk = [bytes([0xc0 + (i % 50)]) + bytes(i % 50) for i in range(200)]
    for n_junk, want_launches in ((128 - len(v["proof_b"]), 1), (129 - len(v["proof_b"]), None)):
        inp = z.MerkleProofInput(v["proof_b"] + junk[:n_junk], v["root_b"], v["key_b"])
        verifier.host_stats(reset=True)
        r = verifier.verify_merkle_proofs([inp])[0]
        assert r == v["value_b"]
        if want_launches:
            assert verifier.host_stats().launches == want_launches
        else:
            assert verifier.host_stats().launches > 1
    for k, one in ((32, True), (33, False)):
        verifier.host_stats(reset=True)
        res = verifier.verify_merkle_proofs([z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"])] * k)
        assert all(r == v["value_b"] for r in res)
        assert (verifier.host_stats().launches == 1) == one
    big = z.MerkleProofInput(v["proof_b"] + [b"\xb9\xff\x00" + bytes(0xff00)], v["root_b"], v["key_b"])  # 65 KB junk node
    verifier.host_stats(reset=True)
    assert verifier.verify_merkle_proofs([big])[0] == v["value_b"]
    assert verifier.host_stats().launches > 1
    assert verifier.verify_merkle_proofs([z.MerkleProofInput([], v["root_b"], b"")])[0].status == 1


def test_storage_guest_flow_is_one_launch(verifier, oracle):
    """account proof + dependent storage proofs (storage-circuit/src/main.rs:6-31) as one small batch: one launch,
    the guest's answers"""
    import zk_state_proofs_b200 as z
    from tests.test_gpu_storage import _guest, _world
    state, accounts = _world(oracle, 11, n_accounts=12, n_slots=12)
    k = oracle.keccak256
    rng = random.Random(12)
    for addr, (slots, st) in accounts.items():
        keys = rng.sample(list(slots), min(3, len(slots)))
        if rng.random() < 0.3:
            keys.append(rng.randbytes(32))
        inp = z.StorageProofInput(state.proof(k(addr)), [st.proof(k(s)) for s in keys], state.root, addr, keys, k(addr))
        want = _guest(oracle, inp)
        verifier.host_stats(reset=True)
        try:
            got = verifier.verify_storage_proof_input(inp)
        except z.VerifyPanic as e:
            got = e.status
        assert got == want
        assert verifier.host_stats().launches == 1  # account + storage proofs + the key hashes: one launch


def test_one_borsh_blob_per_call_is_one_launch(verifier, golden, oracle):
    """mptv_verify_borsh / mptv_verify_storage_borsh with a handful of inputs: flattened on the calling thread and
    verified by ONE launch (no worker pool, no chunk pipeline): the reference's verdict and value for every golden
    vector one blob per call, the guest's outcome for every storage input one blob per call, values reported as
    slices of the blob"""
    import zk_state_proofs_b200 as z
    from tests.test_gpu_storage_borsh import _guest_flow, _inputs
    bad, small_calls = [], 0
    for v in golden["vectors"]:
        blob = z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]).to_borsh()
        verifier.host_stats(reset=True)
        st, voff, vlen = verifier.verify_borsh([blob])
        hs = verifier.host_stats()
        want_st = 6 if len(v["root_b"]) != 32 else v["status"]
        got_v = blob[int(voff[0]):int(voff[0]) + int(vlen[0])] if st[0] == 0 else None
        if (int(st[0]), got_v) != (want_st, v["value_b"] if want_st == 0 else None):
            bad.append((v["tag"], int(st[0]), want_st))
        if len(blob) <= (32 << 10) and len(v["proof_b"]) <= 100 and len(v["root_b"]) == 32:
            small_calls += 1
            assert hs.launches == 1 and hs.chunks <= 1, (v["tag"], hs.launches)
    assert not bad, (len(bad), bad[:10])
    assert small_calls > 2500
    inputs = _inputs(oracle, 123, n_groups=250)
    want = _guest_flow(oracle, inputs)
    one_launch = 0
    for inp, w in zip(inputs, want):
        blob = inp.to_borsh()
        verifier.host_stats(reset=True)
        pf, ist, st, voff, vlen = verifier.verify_storage_borsh([blob])
        hs = verifier.host_stats()
        if isinstance(w, z.VerifyPanic):
            assert int(ist[0]) == w.status
        else:
            assert ist[0] == 0 and [blob[int(voff[q]):int(voff[q]) + int(vlen[q])] for q in range(1, int(pf[1]))] == w
        if len(blob) <= (32 << 10):
            one_launch += hs.launches == 1
    assert one_launch > 200
    # a few inputs in one call still take the path; results equal the one-by-one answers
    blobs = [i.to_borsh() for i in inputs[:6]]
    verifier.host_stats(reset=True)
    pf, ist, st, voff, vlen = verifier.verify_storage_borsh(blobs)
    if sum(len(b) for b in blobs) <= (40 << 10) and int(pf[-1]) <= 32:
        assert verifier.host_stats().launches == 1
    assert [int(x) for x in ist] == [0 if isinstance(w, list) else w.status for w in want[:6]]
