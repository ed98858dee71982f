"""zk-state-proofs_b200 -- B200-native batched Merkle-Patricia-Trie proof verifier.

Host-side mirror of the reference's crypto-ops crate (/root/reference/crypto-ops/src/lib.rs,
keccak.rs, types.rs) over the C ABI of include/mptv.h.  All compute runs in hand-written
sm_100a CUDA (csrc/); there is no CPU fallback -- importing works anywhere, but every compute
call raises if libmptv.so or a B200 is missing.
"""
from .crypto_ops import (  # noqa: F401
    Batch,
    KvBatch,
    MerkleProofInput,
    MptvError,
    StorageProofInput,
    VerifyPanic,
    Verifier,
    STATUS_NAMES,
    digest_keccak,
    Log,
    encode_receipt,
    flatten,
    flatten_borsh,
    flatten_borsh_ex,
    flatten_storage_borsh,
    account_storage_root,
    borsh_flatten_probe,
    host_bw_probe,
    flatten_kv,
    rlp_index_native,
    lib_path,
    load_library,
    ordered_trie_root,
    rlp_index,
    trie_roots,
    verify_merkle_proof,
    verify_merkle_proofs,
    verify_storage_proof_input,
)

__all__ = [
    "Log", "encode_receipt", "flatten_borsh", "flatten_borsh_ex", "flatten_storage_borsh", "account_storage_root", "borsh_flatten_probe", "host_bw_probe", "rlp_index_native",
    "Batch", "KvBatch", "flatten_kv", "ordered_trie_root", "rlp_index", "trie_roots", "MerkleProofInput", "MptvError", "StorageProofInput", "VerifyPanic", "Verifier",
    "STATUS_NAMES", "digest_keccak", "flatten", "lib_path", "load_library", "verify_merkle_proof",
    "verify_merkle_proofs", "verify_storage_proof_input",
]
