"""K4 parity: tx / receipt trie roots rebuilt on the GPU (through the C ABI) against the CPU
restatement of eth_trie's insert + root_hash (oracle/trie_oracle.c) on the same seeded inputs --
bit-exact 32-byte roots -- and closed through the verifier: proofs extracted from the rebuilt trie
verify on the GPU against the GPU-built root and return the inserted bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kv(z, d):
    return z.KvBatch(d["key_bytes"], d["key_off"], d["value_bytes"], d["value_off"], d["value_len"], d["trie_first"])


@pytest.fixture(params=[1, 0], ids=["fused_leaf_hash", "materialised_leaves"])
def leaf_mode(request, verifier):
    """both leaf-level paths: K1L (leaves hashed straight from the value arena) and encode + K1"""
    verifier.set_option("fused_leaf_hash", request.param)
    yield request.param
    verifier.set_option("fused_leaf_hash", 1)


def test_random_tries_roots_match_oracle(verifier, oracle, leaf_mode):
    import zk_state_proofs_b200 as z
    from tests.test_rebuild_oracle import make_kv, random_tries
    for seed in (5, 6, 7):
        tries = random_tries(seed, 300)
        d = make_kv(tries)
        want, perms, hashed = oracle.trie_roots(d, nthreads=8)
        got = verifier.trie_roots(_kv(z, d))
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert len(bad) == 0, [(int(t), len(tries[t])) for t in bad[:10]]


def test_leaf_prefix_alignments_all_phases(verifier, oracle, leaf_mode):
    """every (prefix length mod 16, value length) phase of the fused leaf hash: values of 1..600 bytes under
    keys of 1, 2, 3 and 32 bytes (path items of different lengths), incl. lengths around the 136-byte rate"""
    import zk_state_proofs_b200 as z
    tries = []
    for klen in (1, 2, 3, 32):
        for base in range(0, 600, 50):
            tries.append([(bytes([(7 * i + 1) % 256]) * klen if klen < 32 else bytes([i % 256]) + bytes(31),
                           bytes([(i * 13 + j) % 256 for j in range(base + i + 1)])) for i in range(50)])
    kv = z.flatten_kv(tries)
    assert (verifier.trie_roots(kv) == oracle.trie_roots(kv.as_dict(), nthreads=8)[0]).all()


def test_flatten_kv_and_ordered_trie_root(verifier, oracle):
    import zk_state_proofs_b200 as z
    from oracle.pytrie import Trie
    vals = [b"\x02" + bytes([i % 251]) * (100 + 3 * i) for i in range(200)]
    root = verifier.ordered_trie_root(vals)
    assert root == Trie({z.rlp_index(i): v for i, v in enumerate(vals)}, oracle.keccak256).root
    assert verifier.ordered_trie_root([]).hex() == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"
    # KAT-1 of SURVEY.md Appendix E: single leaf {rlp(0): 0102..08}
    assert verifier.ordered_trie_root([bytes(range(1, 9))]).hex() == \
        "0e9985286c0f4a35519eeb86fa50ce8134ed1fd1bb88e6f74a9e4cb6f505079c"
    assert verifier.trie_roots(z.flatten_kv([])).shape == (0, 32)


def test_config4_shaped_blocks_match_oracle_and_close_through_the_verifier(verifier, oracle, leaf_mode):
    import zk_state_proofs_b200 as z
    from workload import gen
    for kind in ("tx", "receipt"):
        kv = gen.block_tries(150, 300, kind=kind, seed=4)
        want, perms, hashed = oracle.trie_roots(kv.as_dict(), nthreads=8)
        got = verifier.trie_roots(kv)
        assert (got == want).all(), kind
        # timings / work counters of the device entry agree with the oracle's accounting
        # (the host entry ran one chunk on device 0)
        t = verifier.last_rebuild_timings(0)
        assert t.n_perm == perms and t.n_hashed == hashed
        # proofs out of the rebuilt trie (oracle's get_proof) verify on the GPU against the GPU root
        inputs, expect = [], []
        for blk in (0, 77, 149):
            for i in (0, 1, 15, 127, 128, 299):
                key = z.rlp_index(i)
                root, nodes = oracle.trie_get_proof(kv.as_dict(), blk, key)
                assert root == got[blk].tobytes()
                inputs.append(z.MerkleProofInput(nodes, got[blk].tobytes(), key))
                it = blk * 300 + i
                o, ln = int(kv.value_off[it]), int(kv.value_len[it])
                expect.append(kv.value_bytes[o:o + ln].tobytes())
        assert verifier.verify_merkle_proofs(inputs) == expect


def test_refuses_what_it_cannot_hold(verifier, oracle):
    import zk_state_proofs_b200 as z
    with pytest.raises(z.MptvError):
        verifier.trie_roots(z.flatten_kv([[(b"k" * 33, b"v")]]))
    with pytest.raises(z.MptvError):
        verifier.trie_roots(z.flatten_kv([[(i.to_bytes(4, "big"), b"v") for i in range(8193)]]))
    # the largest trie it does hold
    kv = z.flatten_kv([[((i * 2654435761 % (1 << 32)).to_bytes(4, "big"), b"v" * (1 + i % 70)) for i in range(8192)]])
    assert (verifier.trie_roots(kv) == oracle.trie_roots(kv.as_dict())[0]).all()


def _proofs(b):
    return [[b.node_bytes[int(b.node_off[i]):int(b.node_off[i]) + int(b.node_len[i])].tobytes()
             for i in range(int(b.proof_first[q]), int(b.proof_first[q + 1]))] for q in range(b.n_proofs)]


def test_get_proof_matches_oracle_and_verifies(verifier, oracle, leaf_mode):
    """mptv_trie_proofs == eth_trie get_proof as restated by the oracle, node for node, and the batch
    it returns verifies as it is (rebuild -> get_proof -> verify_merkle_proof, all on the GPU)."""
    import random
    import zk_state_proofs_b200 as z
    from tests.test_rebuild_oracle import make_kv, random_tries
    tries = random_tries(9, 120, sizes=(0, 1, 2, 3, 17, 100, 300))
    d = make_kv(tries)
    kv = _kv(z, d)
    rng = random.Random(1)
    targets = []
    for t, kvs in enumerate(tries):
        keys = [k for k, _ in kvs]
        for k in rng.sample(keys, min(3, len(keys))):
            targets.append((t, k))
        targets.append((t, bytes([rng.randrange(256) for _ in range(rng.choice([1, 2, 3, 32]))])))  # mostly absent
    roots, b = verifier.trie_proofs(kv, targets)
    want_roots = oracle.trie_roots(d, nthreads=4)[0]
    assert (roots == want_roots).all()
    got = _proofs(b)
    for (t, k), nodes in zip(targets, got):
        _, want = oracle.trie_get_proof(d, t, k)
        assert nodes == want, (t, k.hex())
    st, voff, vlen = verifier.verify_batch(b)
    ost = [oracle.verify(roots[t].tobytes(), nodes, k)[0] for (t, k), nodes in zip(targets, got)]
    assert st.tolist() == ost
    final = [dict((kk, vv) for kk, vv in kvs) for kvs in tries]
    n_ok = 0
    for q, (t, k) in enumerate(targets):
        v = final[t].get(k, b"")
        if st[q] == 0:
            val = b.value(int(voff[q]), int(vlen[q]))
            assert val == v or (len(v) == 1 and val == b"\x81" + v)
            n_ok += 1
        else:
            # absent / deleted key, or the R4 x R20 quirk on a root-resident value; an EMPTY trie has no
            # node that hashes to its root keccak(0x80), which the reference reports as InvalidStateRoot
            live = any(len(x) for x in final[t].values())
            assert st[q] in (4, 2) if live else st[q] == 1
    assert n_ok > 100


def test_transaction_proof_inputs_roundtrip(verifier, golden):
    """config 1: 200-tx block trie, prove index 15 (trie-utils tests/transaction.rs:13) -- the proof
    built on the GPU is the one in the golden vector produced with the reference ELF."""
    import random
    rng = random.Random(1)
    txs = [b"\x02" + rng.randbytes(rng.randint(99, 299)) for _ in range(200)]
    v = next(v for v in golden["vectors"] if v["tag"] == "config1/tx15")
    inp = verifier.transaction_proof_inputs(txs, 15)
    assert inp.root_hash == v["root_b"] and inp.key == v["key_b"] and inp.proof == v["proof_b"]
    assert verifier.verify_merkle_proof(inp.root_hash, inp.proof, inp.key) == txs[15] == v["value_b"]


def test_rebuild_golden_vectors_on_gpu(verifier, rebuild_golden, leaf_mode):
    """the committed fixture whose proofs were judged by the reference ELF: the GPU rebuild gives the same
    roots, mptv_trie_proofs the same proof nodes, and the GPU verifier the reference's verdicts and values"""
    import zk_state_proofs_b200 as z
    tries = [[(bytes.fromhex(k), bytes.fromhex(v)) for k, v in t["items"]] for t in rebuild_golden["tries"]]
    kv = z.flatten_kv(tries)
    targets, want = [], []
    for t, ent in enumerate(rebuild_golden["tries"]):
        for p in ent["proofs"]:
            targets.append((t, bytes.fromhex(p["key"])))
            want.append(p)
    roots, b = verifier.trie_proofs(kv, targets)
    assert [r.tobytes().hex() for r in roots] == [ent["root"] for ent in rebuild_golden["tries"]]
    got = _proofs(b)
    st, voff, vlen = verifier.verify_batch(b)
    for q, p in enumerate(want):
        assert [x.hex() for x in got[q]] == p["nodes"]
        assert st[q] == p["status"]
        if p["status"] == 0:
            assert b.value(int(voff[q]), int(vlen[q])).hex() == p["value"]


def test_adversarial_shapes_on_gpu(verifier, oracle, leaf_mode):
    """deep prefix chains, 63-nibble extensions, full fan-out, the 32-byte inline boundary, the empty key:
    roots and proofs equal the oracle's, and the proofs verify"""
    import zk_state_proofs_b200 as z
    from tests.test_rebuild_oracle import adversarial_tries, make_kv
    tries = adversarial_tries()
    d = make_kv(tries)
    kv = _kv(z, d)
    want = oracle.trie_roots(d, nthreads=4)[0]
    targets = [(t, k) for t, kvs in enumerate(tries) for k in list(dict(kvs))[:8]]
    targets += [(0, bytes(range(1, 20)) + b"\xee"), (1, b"\x01" * 32), (3, b""), (3, b"\x05\x05")]  # absent keys
    roots, b = verifier.trie_proofs(kv, targets)
    assert (roots == want).all(), np.nonzero((roots != want).any(axis=1))[0]
    got = _proofs(b)
    for (t, k), nodes in zip(targets, got):
        assert nodes == oracle.trie_get_proof(d, t, k)[1], (t, k.hex())
    st, voff, vlen = verifier.verify_batch(b)
    for q, (t, k) in enumerate(targets):
        o = oracle.verify(roots[t].tobytes(), got[q], k)
        assert st[q] == o[0]
        if o[0] == 0:
            assert b.value(int(voff[q]), int(vlen[q])) == o[1]


def test_closed_loop_every_receipt_proves_and_verifies(verifier, leaf_mode):
    """rebuild 60 receipt tries, extract the inclusion proof of EVERY receipt, verify all of them: every verdict
    is OK and every returned value is the inserted receipt (long leaves: up to 30 KB = 221 rate blocks, so the
    long-node launches of K1 and K1L are exercised)"""
    import zk_state_proofs_b200 as z
    from workload import gen
    kv = gen.block_tries(60, 300, "receipt", seed=8)
    keys = [z.rlp_index(i) for i in range(300)]
    targets = [(t, keys[i]) for t in range(60) for i in range(300)]
    roots, b = verifier.trie_proofs(kv, targets)
    st, voff, vlen = verifier.verify_batch(b)
    assert (st == 0).all() and (vlen == kv.value_len).all()
    rng = np.random.default_rng(1)
    for q in rng.choice(len(targets), 400, replace=False):
        o, n = int(kv.value_off[q]), int(kv.value_len[q])
        assert b.value(int(voff[q]), int(vlen[q])) == kv.value_bytes[o:o + n].tobytes()
    # and an absent index proves absent in every trie
    roots2, b2 = verifier.trie_proofs(kv, [(t, z.rlp_index(300)) for t in range(60)])
    assert (roots2 == roots).all() and (verifier.verify_batch(b2)[0] == 4).all()


def test_prefix_heavy_random_tries_on_gpu(verifier, oracle, leaf_mode):
    """2 000 random tries with prefix-sharing keys of 0..32 bytes, overwrites, deletes, values around the inline
    boundary: GPU roots == oracle roots, and sampled proofs == oracle proofs"""
    import zk_state_proofs_b200 as z
    from tests.test_rebuild_oracle import make_kv, prefix_heavy_tries
    for seed in (100, 101, 102, 103, 104):
        tries = prefix_heavy_tries(seed, 400)
        d = make_kv(tries)
        want = oracle.trie_roots(d, nthreads=8)[0]
        targets = [(t, kvs[0][0]) for t, kvs in enumerate(tries)][:150]
        roots, b = verifier.trie_proofs(_kv(z, d), targets)
        bad = np.nonzero((roots != want).any(axis=1))[0]
        assert len(bad) == 0, (seed, [int(t) for t in bad[:5]])
        got = _proofs(b)
        for (t, k), nodes in zip(targets, got):
            assert nodes == oracle.trie_get_proof(d, t, k)[1], (seed, t, k.hex())


def test_receipt_path_closes_from_the_leaf_encoder(verifier, oracle, leaf_mode):
    """the reference's real receipt path end to end (trie-utils/src/receipt.rs:8-38 insert_receipt,
    proofs/receipt.rs:49-92): leaves built by mptv_encode_receipt -- legacy and typed (EIP-2718 prefix), 0 ... 40 logs
    with 0 ... 4 topics and 0 ... 512 bytes of data -- keyed alloy_rlp::encode(index), tries rebuilt on the GPU, the
    proof of EVERY receipt extracted (mptv_trie_proofs) and verified (mptv_verify_batch): each returns exactly the
    encoded receipt, and the roots equal the CPU restatement's"""
    import random
    import zk_state_proofs_b200 as z
    rng = random.Random(77)
    tries, flat = [], []
    for t in range(12):
        n = rng.choice([1, 2, 17, 128, 129, 300])
        leaves = []
        for i in range(n):
            logs = [z.Log(rng.randbytes(20), [rng.randbytes(32) for _ in range(rng.randrange(5))], rng.randbytes(rng.choice([0, 1, 3, 64, 512])))
                    for _ in range(rng.choice([0, 0, 1, 2, 5, 40]))]
            prefix = rng.choice([None, None, 1, 2, 0x7e])
            leaf = z.encode_receipt(rng.random() < 0.9, rng.randrange(1 << rng.choice([8, 24, 40])), rng.randbytes(256), logs, prefix=prefix)
            assert (leaf[0] == prefix) if prefix is not None else leaf[0] >= 0xc0
            leaves.append(leaf)
        tries.append([(z.rlp_index_native(i), v) for i, v in enumerate(leaves)])
        flat += leaves
    kv = z.flatten_kv(tries)
    d = kv.as_dict()
    want_roots = oracle.trie_roots(d, nthreads=4)[0]
    targets = [(t, k) for t, items in enumerate(tries) for k, _ in items]
    roots, b = verifier.trie_proofs(kv, targets)
    assert (roots == want_roots).all()
    st, voff, vlen = verifier.verify_batch(b)
    assert (st == 0).all()
    for q, leaf in enumerate(flat):
        assert b.value(int(voff[q]), int(vlen[q])) == leaf, q
    # the reference-shaped single call on one of them: the Python mirror of get_ethereum_receipt_proof_inputs minus RPC
    inp = verifier.transaction_proof_inputs([v for _, v in tries[5]], 1 if len(tries[5]) > 1 else 0)
    assert verifier.verify_merkle_proof(inp.root_hash, inp.proof, inp.key) == tries[5][1 if len(tries[5]) > 1 else 0][1]
    assert inp.root_hash == want_roots[5].tobytes()
