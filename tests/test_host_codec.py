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
