"""Host-side formats either side of the hot path (csrc/host_codec.cpp), no GPU needed:
the receipt leaf encoder against the reference's own known-answer test
(/root/reference/trie-utils/tests/rlp.rs:12-47), alloy_rlp::encode(index), and the multi-threaded
borsh(MerkleProofInput) -> CSR flattener against the plain Python flatten()."""
import numpy as np
import pytest


def test_receipt_encoding_known_answer_from_the_reference():
    import zk_state_proofs_b200 as z
    expected = bytes.fromhex(
        "f901668001b90100" + "00" * 256 + "f85ff85d940000000000000000000000000000000000000011f842a0"
        "000000000000000000000000000000000000000000000000000000000000deada0"
        "000000000000000000000000000000000000000000000000000000000000beef830100ff")
    logs = [z.Log(bytes.fromhex("0000000000000000000000000000000000000011"),
                  [bytes.fromhex("00" * 30 + "dead"), bytes.fromhex("00" * 30 + "beef")], bytes.fromhex("0100ff"))]
    assert z.encode_receipt(False, 1, bytes(256), logs) == expected
    # typed receipts carry the EIP-2718 type byte in front (receipt.rs:32-35)
    assert z.encode_receipt(False, 1, bytes(256), logs, prefix=2) == b"\x02" + expected
    # status true, big gas, no logs, 1-byte log data < 0x80 (encodes as itself), empty data
    r = z.encode_receipt(True, 0x1234567, b"\xff" * 256, [z.Log(b"\x01" * 20, [], b"\x7f"), z.Log(b"\x02" * 20, [b"\x03" * 32], b"")])
    assert r[:3] == bytes.fromhex("f9015d") and r[3] == 0x01 and r[4:9] == bytes.fromhex("8401234567")
    assert r[9:12] == bytes.fromhex("b90100") and r[12:268] == b"\xff" * 256
    assert r[268:] == bytes.fromhex("f852" + "d7" + "94" + "01" * 20 + "c0" + "7f" + "f8" + "38" + "94" + "02" * 20 + "e1a0" + "03" * 32 + "80")


def test_rlp_index_matches_alloy_encoding():
    import zk_state_proofs_b200 as z
    for i in [0, 1, 15, 0x7f, 0x80, 0xff, 0x100, 299, 0xffff, 0x10000, 2**32 - 1, 2**40 + 5, 2**64 - 1]:
        assert z.rlp_index_native(i) == z.rlp_index(i), i
    assert z.rlp_index(0) == b"\x80" and z.rlp_index(0x80) == b"\x81\x80" and z.rlp_index(300) == b"\x82\x01\x2c"


def test_flatten_borsh_equals_python_flatten(golden):
    import zk_state_proofs_b200 as z
    inputs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]) for v in golden["vectors"]]
    inputs += [z.MerkleProofInput([], b"", b""), z.MerkleProofInput([b""], b"\x01" * 31, b"\x05" * 40)]
    blobs = [i.to_borsh() for i in inputs]
    want = z.flatten(inputs)
    for threads in (1, 4, 0):
        got = z.flatten_borsh(blobs, threads=threads)
        assert got.n_proofs == want.n_proofs and got.n_nodes == want.n_nodes
        for name in ["node_off", "node_len", "proof_first", "roots", "key_off"]:
            assert (getattr(got, name) == getattr(want, name)).all(), name
        assert (got.node_bytes == want.node_bytes).all()
        assert (got.key_bytes[:int(got.key_off[-1])] == want.key_bytes[:int(want.key_off[-1])]).all()
        assert (got.bad_root_len == want.bad_root_len).all()
    assert z.flatten_borsh([]).n_proofs == 0


def test_flatten_borsh_rejects_what_borsh_rejects():
    import zk_state_proofs_b200 as z
    good = z.MerkleProofInput([b"\x01\x02"], b"\xaa" * 32, b"\x07").to_borsh()
    for bad in [good[:-1], good + b"\x00", good[:3], b"\xff\xff\xff\xff" + good[4:], b""]:
        with pytest.raises(ValueError):
            z.flatten_borsh([good, bad])


def test_property_borsh_roundtrip_and_flatten_agree():
    """hypothesis: arbitrary MerkleProofInputs survive borsh, and the C++ flattener lays them out exactly like the
    plain Python flatten() (offsets, lengths, 16-byte alignment, roots, keys, bad-root flags)"""
    from hypothesis import given, settings, strategies as st
    import zk_state_proofs_b200 as z

    node = st.binary(min_size=0, max_size=600)
    inp = st.builds(z.MerkleProofInput, st.lists(node, min_size=0, max_size=9),
                    st.one_of(st.binary(min_size=32, max_size=32), st.binary(min_size=0, max_size=40)),
                    st.binary(min_size=0, max_size=70))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(inp, min_size=0, max_size=12))
    def check(inputs):
        blobs = [i.to_borsh() for i in inputs]
        assert [z.MerkleProofInput.from_borsh(b) for b in blobs] == inputs
        got, want = z.flatten_borsh(blobs, threads=2), z.flatten(inputs)
        assert got.n_proofs == want.n_proofs and got.n_nodes == want.n_nodes
        for name in ["node_off", "node_len", "proof_first", "roots", "key_off"]:
            assert (getattr(got, name) == getattr(want, name)).all()
        for o, n in zip(want.node_off.tolist(), want.node_len.tolist()):
            assert (got.node_bytes[o:o + n] == want.node_bytes[o:o + n]).all()
        k = int(want.key_off[-1])
        assert (got.key_bytes[:k] == want.key_bytes[:k]).all()
        gb = np.zeros(len(inputs), bool) if got.bad_root_len is None else got.bad_root_len
        wb = np.zeros(len(inputs), bool) if want.bad_root_len is None else want.bad_root_len
        assert (gb == wb).all()

    check()


def test_host_pool_nt_copy_and_borsh_walkers_cpp(tmp_path):
    """tests/cpp/test_host_pool.cpp: the worker pool + barrier behind mptv_verify_borsh (also under
    -fsanitize=thread when the toolchain has it), the non-temporal node copy, the borsh walkers."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "test_host_pool.cpp")
    exe = str(tmp_path / "test_host_pool")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    tsan = str(tmp_path / "test_host_pool_tsan")
    if subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", src, "-o", tsan],
                      capture_output=True).returncode == 0:
        r = subprocess.run([tsan], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stdout + r.stderr


def _node_view(b, k):
    o, n = int(b.node_off[k]), int(b.node_len[k])
    return b.node_bytes[o:o + n].tobytes()


def test_flatten_borsh_ex_single_pass_and_aliasing(golden, oracle):
    """mptv_flatten_borsh_ex (csrc/host_flatten.h): the one-pass builder reproduces every node, root and key of the
    blobs; with MPTV_FLATTEN_ALIAS_DUPLICATES byte-identical nodes share one copy, and the oracle gives the same
    verdicts and values on the aliased arena as on the plain one."""
    import zk_state_proofs_b200 as z
    inputs = [z.MerkleProofInput(v["proof_b"], v["root_b"], v["key_b"]) for v in golden["vectors"]]
    inputs += [z.MerkleProofInput([], b"", b""), z.MerkleProofInput([b""], b"\x01" * 31, b"\x05" * 40)]
    # many proofs against one trie: the shared upper nodes (>= 128 bytes) are what aliasing removes
    from oracle.pytrie import Trie
    import random
    rng = random.Random(5)
    kv = {oracle.keccak256(bytes([i, j])): rng.randbytes(70) for i in range(40) for j in range(40)}
    t = Trie(kv, oracle.keccak256)
    keys = list(kv)
    inputs += [z.MerkleProofInput(t.proof(k), t.root, k) for k in keys[:600]]
    blobs = [i.to_borsh() for i in inputs]
    want = z.flatten(inputs)
    ref = oracle.verify_batch(dict(node_bytes=want.node_bytes, node_off=want.node_off, node_len=want.node_len,
                                   proof_first=want.proof_first, roots=want.roots, key_bytes=want.key_bytes,
                                   key_off=want.key_off), nthreads=2)
    for alias in (False, True):
        for threads in (1, 3, 0):
            got, info = z.flatten_borsh_ex(blobs, threads=threads, alias_duplicates=alias)
            assert got.n_proofs == want.n_proofs and got.n_nodes == want.n_nodes == info.n_nodes
            assert (got.node_len == want.node_len).all() and (got.proof_first == want.proof_first).all()
            assert (got.roots == want.roots).all() and (got.key_off == want.key_off).all()
            assert (got.node_off % 16 == 0).all()
            k = int(want.key_off[-1])
            assert (got.key_bytes[:k] == want.key_bytes[:k]).all()
            assert ((got.bad_root_len if got.bad_root_len is not None else np.zeros(len(inputs), bool)) ==
                    (want.bad_root_len if want.bad_root_len is not None else np.zeros(len(inputs), bool))).all()
            for kk in range(want.n_nodes):
                assert _node_view(got, kk) == _node_view(want, kk), kk
            distinct_offsets = len(set(got.node_off.tolist()))
            if alias:
                assert info.nodes_aliased > 1000 and info.node_bytes_placed < info.node_bytes_supplied - 128 * info.nodes_aliased + 1
                assert distinct_offsets <= want.n_nodes - info.nodes_aliased + 1
            else:
                assert info.nodes_aliased == 0 and info.node_bytes_placed == info.node_bytes_supplied
            r = oracle.verify_batch(dict(node_bytes=got.node_bytes, node_off=got.node_off, node_len=got.node_len,
                                         proof_first=got.proof_first, roots=got.roots, key_bytes=got.key_bytes,
                                         key_off=got.key_off), nthreads=2)
            assert (r[0] == ref[0]).all() and (r[2] == ref[2]).all()
            for i in np.nonzero(r[0] == 0)[0]:
                assert got.value(int(r[1][i]), int(r[2][i])) == want.value(int(ref[1][i]), int(ref[2][i]))


def test_flatten_ex_never_aliases_nodes_that_differ_in_one_byte():
    """the fingerprint only selects candidates: nodes that agree on every sampled word but differ elsewhere stay apart"""
    import zk_state_proofs_b200 as z
    base = bytearray(b"\xf9\x02\x11" + b"\xa0" + bytes(range(200)) * 3)[:532]
    nodes = []
    for pos in range(0, 532, 1):
        n = bytearray(base)
        n[pos] ^= 0x40
        nodes.append(bytes(n))
    nodes += nodes[:50] + [bytes(base)] * 3
    inputs = [z.MerkleProofInput(nodes[i:i + 7], b"\xaa" * 32, b"\x01") for i in range(0, len(nodes), 7)]
    got, info = z.flatten_borsh_ex([i.to_borsh() for i in inputs], threads=2, alias_duplicates=True)
    flat = [n for i in inputs for n in i.proof]
    # 404 of the 532 variants agree on every sampled word: they chain in one bucket and only the first few (the probe
    # budget) are recorded, so not every later duplicate is shared -- but nothing that differs is ever merged
    assert 3 <= info.nodes_aliased <= 53
    for k, n in enumerate(flat):
        assert _node_view(got, k) == n


def test_flatten_ex_rejects_what_borsh_rejects_and_probe_runs():
    import zk_state_proofs_b200 as z
    good = z.MerkleProofInput([b"\x01\x02" * 100], b"\xaa" * 32, b"\x07").to_borsh()
    for bad in [good[:-1], good + b"\x00", good[:3], b"\xff\xff\xff\xff" + good[4:], b"", good[:11]]:
        with pytest.raises(ValueError):
            z.flatten_borsh_ex([good, bad])
        with pytest.raises(ValueError):
            z.borsh_flatten_probe([good, bad])
    dt, info = z.borsh_flatten_probe([good] * 1000, threads=2, chunk_bytes=1 << 14)
    assert info.n_nodes == 1000 and info.nodes_aliased > 900 and dt > 0
    b, info = z.flatten_borsh_ex([])
    assert b.n_proofs == 0


def test_storage_borsh_flattener_matches_the_structs(oracle):
    """mptv_flatten_storage_borsh (the host half of mptv_verify_storage_borsh): every StorageProofInput becomes its
    account proof + min(#storage_proofs, #storage_keys) storage proofs with raw keys, hash flags and root_from_proof
    (crypto-ops/src/types.rs:11-19, storage-circuit/src/main.rs:10-27), with and without aliasing, 1 and 3 threads"""
    import zk_state_proofs_b200 as z
    from tests.test_gpu_storage_borsh import _inputs
    inputs = _inputs(oracle, 11, 300)
    blobs = [i.to_borsh() for i in inputs]
    for alias in (False, True):
        for th in (1, 3):
            b, hk, pf, info = z.flatten_storage_borsh(blobs, threads=th, alias_duplicates=alias)
            q = 0
            for i, inp in enumerate(inputs):
                used = min(len(inp.storage_proofs), len(inp.storage_keys))
                assert int(pf[i]) == q and int(pf[i + 1]) == q + 1 + used
                items = [(inp.account_proof, inp.address_keccak, -1, 0)] + \
                        [(inp.storage_proofs[j], inp.storage_keys[j], q, 1) for j in range(used)]
                for j, (pr, key, r, h) in enumerate(items):
                    p = q + j
                    f, e = int(b.proof_first[p]), int(b.proof_first[p + 1])
                    assert e - f == len(pr)
                    for k, nd in enumerate(pr):
                        no, nl = int(b.node_off[f + k]), int(b.node_len[f + k])
                        assert bytes(b.node_bytes[no:no + nl]) == nd
                    assert bytes(b.key_bytes[int(b.key_off[p]):int(b.key_off[p + 1])]) == bytes(key)
                    assert int(b.root_from_proof[p]) == r and int(hk[p]) == h
                    assert bytes(b.roots[32 * p:32 * p + 32]) == (inp.root_hash if (j == 0 and len(inp.root_hash) == 32) else bytes(32))
                assert bool(b.bad_root_len is not None and b.bad_root_len[q]) == (len(inp.root_hash) != 32)
                q += 1 + used
            assert (info.nodes_aliased > 0) == alias and info.n_nodes == sum(len(p) for p in [i.account_proof for i in inputs]) + \
                sum(len(pr) for i in inputs for pr in i.storage_proofs[:min(len(i.storage_proofs), len(i.storage_keys))])
    good = blobs[0]
    for bad in (good[:-1], good + b"\0", good[:3], b"", b"\xff\xff\xff\xff" + good[4:], good[:-33]):
        with pytest.raises(ValueError):
            z.flatten_storage_borsh(blobs[:5] + [bad] + blobs[5:])


def test_account_decode_on_the_host_matches_the_oracle(oracle):
    """mptv_account_storage_root = alloy_rlp::decode_exact::<Account> of the storage guest (main.rs:15), the rule the
    device applies too (oracle: mpto_account_storage_root)"""
    import random
    import zk_state_proofs_b200 as z
    from oracle.pytrie import rlp_list, rlp_str, rlp_uint
    rng = random.Random(1)
    ok = 0
    for _ in range(5000):
        k = rng.random()
        items = [rlp_uint(rng.randrange(0, 1 << rng.choice([0, 7, 8, 63, 64, 70]))),
                 rlp_uint(rng.randrange(0, 1 << rng.choice([0, 8, 255, 256, 260]))),
                 rlp_str(rng.randbytes(rng.choice([32, 32, 32, 31, 33, 0]))), rlp_str(rng.randbytes(rng.choice([32, 32, 32, 31])))]
        if k < 0.1:
            items = items[:3]
        elif k < 0.2:
            items.append(b"\x80")
        elif k < 0.25:
            items[1] = rlp_str(b"\x00" + rng.randbytes(3))  # leading zero
        elif k < 0.3:
            items[0] = b"\x00"
        v = rlp_list(items)
        if k > 0.95:
            v += b"\x00"
        elif k > 0.9:
            v = v[:-1]
        got = z.account_storage_root(v)
        assert got == oracle.account_storage_root(v)
        ok += got is not None
    assert 500 < ok < 4500 and z.account_storage_root(b"") is None and z.account_storage_root(b"\xc0") is None


def test_storage_borsh_parser_accepts_exactly_what_the_mirror_accepts(oracle):
    """untrusted bytes: valid StorageProofInput blobs with corrupted length words, truncations, trailing bytes and bit
    flips -- mptv_flatten_storage_borsh accepts a blob iff the Python mirror's from_borsh (borsh::from_slice semantics)
    does, lays it out identically when it does, and never reads outside the blob"""
    import random
    import struct
    import zk_state_proofs_b200 as z
    from tests.test_gpu_storage_borsh import _inputs
    blobs = [i.to_borsh() for i in _inputs(oracle, 31, 120)]
    rng = random.Random(7)
    n_acc = n_rej = 0
    for it in range(2000):
        b = bytearray(blobs[rng.randrange(len(blobs))])
        k = rng.random()
        if k < 0.3:
            p = rng.randrange(0, len(b) - 4)
            struct.pack_into("<I", b, p, rng.choice([0, 1, 2, 3, 5, 0xffffffff, 0x7fffffff, len(b), len(b) - p, rng.randrange(0, 1 << 16)]))
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
            want = z.StorageProofInput.from_borsh(b)
        except Exception:
            want = None
        try:
            fb, hk, pf, info = z.flatten_storage_borsh([blobs[0], b, blobs[1]])
        except ValueError:
            assert want is None, (it, len(b))
            n_rej += 1
            continue
        assert want is not None, (it, len(b))
        n_acc += 1
        used = min(len(want.storage_proofs), len(want.storage_keys))
        q = int(pf[1])
        assert int(pf[2]) - q == 1 + used
        for j, pr in enumerate([want.account_proof] + want.storage_proofs[:used]):
            f, e = int(fb.proof_first[q + j]), int(fb.proof_first[q + j + 1])
            assert e - f == len(pr)
            for kx, nd in enumerate(pr):
                no, nl = int(fb.node_off[f + kx]), int(fb.node_len[f + kx])
                assert bytes(fb.node_bytes[no:no + nl]) == nd
        for j in range(used):
            assert bytes(fb.key_bytes[int(fb.key_off[q + 1 + j]):int(fb.key_off[q + 2 + j])]) == bytes(want.storage_keys[j])
    assert n_acc > 300 and n_rej > 300
