"""StorageProofInput (crypto-ops/src/types.rs:11-19) through the batched GPU path against the storage
guest's flow (storage-circuit/src/main.rs:6-31) restated with the oracle: verify the account proof under
address_keccak, decode the Account RLP, verify every storage proof under its storage_root with key
keccak(storage_key)."""
import random

import pytest

pytestmark = pytest.mark.gpu


def _rlp_account(nonce, balance, storage_root, code_hash):
    from oracle.pytrie import rlp_list, rlp_str, rlp_uint
    return rlp_list([rlp_uint(nonce), rlp_uint(balance), rlp_str(storage_root), rlp_str(code_hash)])


def _world(oracle, seed, n_accounts=40, n_slots=30):
    from oracle.pytrie import Trie, rlp_uint
    rng = random.Random(seed)
    k = oracle.keccak256
    storages, accounts = [], {}
    for a in range(n_accounts):
        addr = rng.randbytes(20)
        slots = {rng.randbytes(32): rng.randrange(1, 1 << rng.choice([7, 8, 64, 255])) for _ in range(rng.randrange(1, n_slots))}
        st = Trie({k(s): rlp_uint(v) for s, v in slots.items()}, k)
        accounts[addr] = (slots, st)
        storages.append(st)
    state = Trie({k(addr): _rlp_account(rng.randrange(1 << 20), rng.randrange(1 << 100), st.root, rng.randbytes(32))
                  for addr, (slots, st) in accounts.items()}, k)
    return state, accounts


def _guest(oracle, inp):
    """the storage guest, with the oracle as verify_merkle_proof"""
    st, val, _, _ = oracle.verify(inp.root_hash, inp.account_proof, inp.address_keccak)
    if st != 0:
        return st
    sr = oracle.account_storage_root(val)
    if sr is None:
        return 7
    out = []
    for pr, key in zip(inp.storage_proofs, inp.storage_keys):
        st, val, _, _ = oracle.verify(sr, pr, oracle.keccak256(key))
        if st != 0:
            return st
        out.append(val)
    return out


def test_storage_proof_inputs_batched_match_guest_flow(verifier, oracle):
    import zk_state_proofs_b200 as z
    state, accounts = _world(oracle, 3)
    rng = random.Random(4)
    k = oracle.keccak256
    inputs = []
    for addr, (slots, st) in accounts.items():
        keys = rng.sample(list(slots), min(3, len(slots)))
        if rng.random() < 0.3:
            keys.append(rng.randbytes(32))  # absent slot -> "Key does not exist!"
        inp = z.StorageProofInput(state.proof(k(addr)), [st.proof(k(s)) for s in keys], state.root, addr, keys, k(addr))
        u = rng.random()
        if u < 0.1:
            inp.account_proof = inp.account_proof[:-1]            # truncated account proof
        elif u < 0.2 and inp.storage_proofs[0]:
            n = bytearray(inp.storage_proofs[0][-1]); n[-1] ^= 1
            inp.storage_proofs[0] = inp.storage_proofs[0][:-1] + [bytes(n)]   # tampered storage leaf
        elif u < 0.3:
            inp.address_keccak = k(b"someone else")               # wrong account key
        inputs.append(inp)
    got = verifier.verify_storage_proof_inputs(inputs)
    n_ok = 0
    for inp, g in zip(inputs, got):
        want = _guest(oracle, inp)
        if isinstance(want, int):
            assert isinstance(g, z.VerifyPanic) and g.status == want, (g, want)
        else:
            assert g == want
            n_ok += 1
            assert verifier.verify_storage_proof_input(inp) == want   # the single-input entry agrees
    assert n_ok >= 15 and n_ok < len(inputs)


def test_hashed_keys_entry_equals_two_step_flow(verifier, oracle):
    """mptv_verify_batch_hashed_keys (storage keys hashed on the device inside the call) == hashing them first"""
    import numpy as np
    import zk_state_proofs_b200 as z
    state, accounts = _world(oracle, 9, n_accounts=25)
    k = oracle.keccak256
    items, rfp, flags, items_hashed = [], [], [], []
    for addr, (slots, st) in accounts.items():
        a = len(items)
        items.append(z.MerkleProofInput(state.proof(k(addr)), state.root, k(addr)))
        items_hashed.append(items[-1])
        rfp.append(-1)
        flags.append(0)
        for s in list(slots)[:3] + [b"\x07" * 32]:
            items.append(z.MerkleProofInput(st.proof(k(s)), b"\x00" * 32, s))            # raw slot key
            items_hashed.append(z.MerkleProofInput(st.proof(k(s)), b"\x00" * 32, k(s)))  # hashed by the caller
            rfp.append(a)
            flags.append(1)
    got = verifier.verify_batch_hashed_keys(z.flatten(items, rfp), np.array(flags, np.uint8))
    want = verifier.verify_batch(z.flatten(items_hashed, rfp))
    for x, y in zip(got, want):
        assert (x == y).all()
    assert (got[0] == 0).sum() > 60 and (got[0] == 4).sum() >= 20


def test_hashed_keys_on_the_device_all_key_lengths_both_paths(verifier, oracle):
    """k_prepare_keys (batch pipeline) and the latency kernel's key step hash the flagged keys on the device, straight
    from the packed key arena: keys of 0 ... 300 bytes (every rate-block boundary case), flagged and unflagged mixed,
    small batches (one launch) and large ones (chunked) give what hashing on the host first gives"""
    import numpy as np
    import zk_state_proofs_b200 as z
    from oracle.pytrie import Trie
    k = oracle.keccak256
    rng = random.Random(21)
    lens = [0, 1, 31, 32, 33, 64, 134, 135, 136, 137, 271, 272, 273, 300]
    raw = [rng.randbytes(n) for n in lens for _ in range(3)]
    t = Trie({k(r): rng.randbytes(rng.randrange(1, 40)) for r in raw}, k)
    for n_items, chunk in ((7, 96 << 20), (len(raw), 96 << 20), (len(raw), 1 << 16)):
        items, flags, want = [], [], []
        for i, r in enumerate(raw[:n_items]):
            if i % 3 == 2:   # unflagged: the caller hashed already
                items.append(z.MerkleProofInput(t.proof(k(r)), t.root, k(r))); flags.append(0)
            else:
                items.append(z.MerkleProofInput(t.proof(k(r)), t.root, r)); flags.append(1)
            want.append(oracle.verify(t.root, t.proof(k(r)), k(r))[:2])
        b = z.flatten(items)
        verifier.set_option("chunk_bytes", chunk)
        try:
            for lp in (1, 0):
                verifier.set_option("latency_path", lp)
                st, voff, vlen = verifier.verify_batch_hashed_keys(b, np.array(flags, np.uint8))
                for p, (ws, wv) in enumerate(want):
                    assert st[p] == ws == 0 and b.value(int(voff[p]), int(vlen[p])) == wv, (n_items, chunk, lp, p)
                # without the flags the raw keys are not what the trie holds (absent, or into an unproven subtree)
                st2, _, _ = verifier.verify_batch(b)
                assert all(st2[p] == 0 for p in range(n_items) if not flags[p])
                assert all(st2[p] in (3, 4) for p in range(n_items) if flags[p])
        finally:
            verifier.set_option("chunk_bytes", 96 << 20)
            verifier.set_option("latency_path", 1)
