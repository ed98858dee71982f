// mptv_crypto_ops.hpp -- C++ host-side mirror of the reference's crypto-ops / trie-utils interface
// over the C ABI of mptv.h (header-only; link libmptv.so).
//
// The reference's host side is Rust and this image has no Rust toolchain, so the reference-shaped
// host API is provided in C++ with the same names, argument meaning and error behaviour:
//
//   crypto_ops::verify_merkle_proof(root_hash, proof, key) -> Bytes     crypto-ops/src/lib.rs:8-23
//       the reference panics on failure; here a crypto_ops::VerifyPanic is thrown whose what() is
//       the reference's panic message class and whose status is the MPTV_ST_* verdict
//   crypto_ops::verify_merkle_proofs(inputs) -> vector<Outcome>          (the batched entry)
//   crypto_ops::digest_keccak(bytes) -> B256                             crypto-ops/src/keccak.rs:6-12
//   crypto_ops::MerkleProofInput / StorageProofInput (+ borsh)           crypto-ops/src/types.rs:4-19
//   crypto_ops::verify_storage_proof_input(input) -> vector<Bytes>       storage-circuit/src/main.rs:6-31
//   trie_utils::Log, encode_receipt, rlp_index                           trie-utils/src/types.rs:11-35, receipt.rs:8-38
//   trie_utils::ordered_trie_root, transaction_proof_inputs              trie-utils/src/proofs/transaction.rs:41-73
//
// There is no CPU fallback: constructing a Verifier without a B200 throws.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mptv.h"

namespace crypto_ops {

using Bytes = std::vector<uint8_t>;
using B256 = std::array<uint8_t, 32>;

inline void borsh_put_u32(Bytes& o, uint32_t v) { for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i))); }
inline void borsh_put_bytes(Bytes& o, const Bytes& b) { borsh_put_u32(o, (uint32_t)b.size()); o.insert(o.end(), b.begin(), b.end()); }
struct BorshReader {
  const uint8_t* p; const uint8_t* end;
  uint32_t u32() {
    if (end - p < 4) throw std::invalid_argument("borsh: unexpected end of input");
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    p += 4;
    return v;
  }
  Bytes bytes() {
    uint32_t n = u32();
    if ((uint64_t)(end - p) < n) throw std::invalid_argument("borsh: unexpected end of input");
    Bytes b(p, p + n);
    p += n;
    return b;
  }
};

// crypto-ops/src/types.rs:4-9
struct MerkleProofInput {
  std::vector<Bytes> proof;
  Bytes root_hash;
  Bytes key;
  Bytes to_borsh() const {
    Bytes o;
    borsh_put_u32(o, (uint32_t)proof.size());
    for (const Bytes& n : proof) borsh_put_bytes(o, n);
    borsh_put_bytes(o, root_hash);
    borsh_put_bytes(o, key);
    return o;
  }
  static MerkleProofInput from_borsh(const uint8_t* data, size_t len) {
    BorshReader r{data, data + len};
    MerkleProofInput m;
    uint32_t n = r.u32();
    for (uint32_t i = 0; i < n; i++) m.proof.push_back(r.bytes());
    m.root_hash = r.bytes();
    m.key = r.bytes();
    if (r.p != r.end) throw std::invalid_argument("borsh: trailing bytes after MerkleProofInput");
    return m;
  }
  bool operator==(const MerkleProofInput& o) const { return proof == o.proof && root_hash == o.root_hash && key == o.key; }
};

// crypto-ops/src/types.rs:11-19 (storage keys are un-hashed; the consumer hashes them)
struct StorageProofInput {
  std::vector<Bytes> account_proof;
  std::vector<std::vector<Bytes>> storage_proofs;
  Bytes root_hash;
  Bytes account_key;
  std::vector<Bytes> storage_keys;
  B256 address_keccak{};
  // borsh derive order = field order; the fixed-size array has no length prefix
  Bytes to_borsh() const {
    Bytes o;
    borsh_put_u32(o, (uint32_t)account_proof.size());
    for (const Bytes& n : account_proof) borsh_put_bytes(o, n);
    borsh_put_u32(o, (uint32_t)storage_proofs.size());
    for (const std::vector<Bytes>& pr : storage_proofs) {
      borsh_put_u32(o, (uint32_t)pr.size());
      for (const Bytes& n : pr) borsh_put_bytes(o, n);
    }
    borsh_put_bytes(o, root_hash);
    borsh_put_bytes(o, account_key);
    borsh_put_u32(o, (uint32_t)storage_keys.size());
    for (const Bytes& k : storage_keys) borsh_put_bytes(o, k);
    o.insert(o.end(), address_keccak.begin(), address_keccak.end());
    return o;
  }
  static StorageProofInput from_borsh(const uint8_t* data, size_t len) {
    BorshReader r{data, data + len};
    StorageProofInput s;
    for (uint32_t n = r.u32(), i = 0; i < n; i++) s.account_proof.push_back(r.bytes());
    for (uint32_t m = r.u32(), j = 0; j < m; j++) {
      s.storage_proofs.emplace_back();
      for (uint32_t n = r.u32(), i = 0; i < n; i++) s.storage_proofs.back().push_back(r.bytes());
    }
    s.root_hash = r.bytes();
    s.account_key = r.bytes();
    for (uint32_t k = r.u32(), i = 0; i < k; i++) s.storage_keys.push_back(r.bytes());
    if (r.end - r.p != 32) throw std::invalid_argument("borsh: StorageProofInput must end with the 32-byte address_keccak");
    memcpy(s.address_keccak.data(), r.p, 32);
    return s;
  }
  bool operator==(const StorageProofInput& o) const {
    return account_proof == o.account_proof && storage_proofs == o.storage_proofs && root_hash == o.root_hash &&
           account_key == o.account_key && storage_keys == o.storage_keys && address_keccak == o.address_keccak;
  }
};

// what the reference would have panicked with
struct VerifyPanic : std::runtime_error {
  int status;
  explicit VerifyPanic(int st) : std::runtime_error(message(st)), status(st) {}
  static const char* message(int st) {
    switch (st) {
      case MPTV_ST_INVALID_STATE_ROOT: return "Invalid merkle proof: InvalidStateRoot";
      case MPTV_ST_ROOT_NOT_CANONICAL: return "assertion `left == right` failed";
      case MPTV_ST_INVALID_PROOF: return "Failed to verify Merkle Proof: InvalidProof";
      case MPTV_ST_KEY_NOT_FOUND: return "Key does not exist!";
      case MPTV_ST_BAD_ROOT_LEN: return "called `Result::unwrap()` on an `Err` value: TryFromSliceError";
      case MPTV_ST_DEP_FAILED: return "account proof rejected or not an Account RLP";
      default: return "panicked inside eth_trie (invalid data / index out of bounds)";
    }
  }
};
struct MptvError : std::runtime_error { using std::runtime_error::runtime_error; };

// per-proof outcome of the batched entry: the value, or the verdict the reference would panic with
struct Outcome {
  int status = 0;
  Bytes value;
  bool ok() const { return status == MPTV_ST_OK; }
};

class Verifier {
 public:
  Verifier() : Verifier(std::vector<int>{}) {}
  explicit Verifier(const std::vector<int>& device_ids) {
    int rc = mptv_create(device_ids.empty() ? nullptr : device_ids.data(), (int)device_ids.size(), &ctx_);
    if (rc != MPTV_OK) throw MptvError(std::string("mptv_create: ") + mptv_strerror(rc));
  }
  ~Verifier() {
    mptv_host_batch_free(pinned_);
    mptv_host_batch_free(pageable_);
    mptv_destroy(ctx_);
  }
  Verifier(const Verifier&) = delete;
  Verifier& operator=(const Verifier&) = delete;
  mptv_ctx* ctx() { return ctx_; }

  // batched entry: flatten (borsh -> CSR through the multi-threaded C++ flattener), verify, slice
  std::vector<Outcome> verify_merkle_proofs(const std::vector<MerkleProofInput>& inputs,
                                            const std::vector<int32_t>* root_from_proof = nullptr) {
    std::vector<Outcome> out(inputs.size());
    if (inputs.empty()) return out;
    Bytes blobs;
    std::vector<uint64_t> off(inputs.size() + 1, 0);
    for (size_t i = 0; i < inputs.size(); i++) {
      Bytes b = inputs[i].to_borsh();
      blobs.insert(blobs.end(), b.begin(), b.end());
      off[i + 1] = blobs.size();
    }
    // Large plain batches go through the streamed entry: blobs in, verdicts out, the flattening of one chunk
    // overlapping the copy and the kernels of the previous ones
    if (!root_from_proof && blobs.size() > (8u << 20)) return verify_borsh(blobs.data(), off.data(), inputs.size());
    // Large batches are flattened straight into PAGE-LOCKED buffers (kept and recycled across calls): the
    // host-buffer entry then streams them at PCIe speed (a copy from pageable memory runs ~5x slower).
    // Small ones stay pageable: mptv_verify_batch packs them into its own pinned staging block anyway.
    const bool big = blobs.size() > (1u << 20);
    mptv_host_batch*& hb = big ? pinned_ : pageable_;
    int rc = mptv_flatten_borsh(blobs.data(), off.data(), inputs.size(), 0, big ? 1 : 0, &hb);
    if (rc == MPTV_ERR_NOMEM) hb = nullptr;  // the flattener released the handle
    if (rc != MPTV_OK) throw MptvError(std::string("mptv_flatten_borsh: ") + mptv_strerror(rc));
    mptv_batch b = *mptv_host_batch_view(hb);
    if (root_from_proof) b.root_from_proof = root_from_proof->data();
    const size_t n = inputs.size();
    std::vector<uint8_t> status(n);
    std::vector<uint64_t> voff(n);
    std::vector<uint32_t> vlen(n);
    mptv_result r{status.data(), voff.data(), vlen.data()};
    check(mptv_verify_batch(ctx_, &b, &r), "mptv_verify_batch");
    const uint8_t* bad = mptv_host_batch_bad_root(hb);
    for (size_t p = 0; p < n; p++) {
      out[p].status = bad[p] ? MPTV_ST_BAD_ROOT_LEN : status[p];
      if (out[p].ok()) out[p].value.assign(b.node_bytes + voff[p], b.node_bytes + voff[p] + vlen[p]);
    }
    return out;
  }

  // borsh(MerkleProofInput) blobs (blob i = blobs[off[i] .. off[i+1])) -> outcomes, through mptv_verify_borsh
  std::vector<Outcome> verify_borsh(const uint8_t* blobs, const uint64_t* off, size_t n) {
    std::vector<Outcome> out(n);
    if (n == 0) return out;
    std::vector<uint8_t> status(n);
    std::vector<uint64_t> voff(n);
    std::vector<uint32_t> vlen(n);
    mptv_result r{status.data(), voff.data(), vlen.data()};
    check(mptv_verify_borsh(ctx_, blobs, off, n, 0, &r), "mptv_verify_borsh");
    for (size_t p = 0; p < n; p++) {
      out[p].status = status[p];
      if (out[p].ok()) out[p].value.assign(blobs + voff[p], blobs + voff[p] + vlen[p]);  // a slice of the blob itself
    }
    return out;
  }

  // crypto-ops/src/lib.rs:8 -- takes ownership of proof, borrows key, returns the value or "panics"
  Bytes verify_merkle_proof(const B256& root_hash, std::vector<Bytes> proof, const Bytes& key) {
    MerkleProofInput in{std::move(proof), Bytes(root_hash.begin(), root_hash.end()), key};
    Outcome o = verify_merkle_proofs({in})[0];
    if (!o.ok()) throw VerifyPanic(o.status);
    return o.value;
  }

  // crypto-ops/src/keccak.rs:6
  B256 digest_keccak(const uint8_t* data, size_t len) { return keccak_many({Bytes(data, data + len)})[0]; }
  B256 digest_keccak(const Bytes& b) { return digest_keccak(b.data(), b.size()); }

  std::vector<B256> keccak_many(const std::vector<Bytes>& msgs) {
    std::vector<B256> out(msgs.size());
    if (msgs.empty()) return out;
    Bytes arena;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len;
    for (const Bytes& m : msgs) {
      off.push_back(arena.size());
      len.push_back((uint32_t)m.size());
      arena.insert(arena.end(), m.begin(), m.end());
      arena.resize((arena.size() + 15) & ~(size_t)15, 0);
    }
    arena.resize(arena.size() + 16, 0);
    check(mptv_keccak256_batch(ctx_, arena.data(), arena.size(), off.data(), len.data(), msgs.size(), out[0].data()),
          "mptv_keccak256_batch");
    return out;
  }

  // the risc0 storage guest: account proof under address_keccak, then every storage proof under the
  // account's storage_root with key keccak(storage_key); returns the storage values, throws like the guest
  std::vector<Bytes> verify_storage_proof_input(const StorageProofInput& in) {
    // the guest zips storage_proofs with storage_keys (main.rs:18-21): the shorter list decides
    const size_t n_pairs = std::min(in.storage_proofs.size(), in.storage_keys.size());
    std::vector<B256> hashed = keccak_many(std::vector<Bytes>(in.storage_keys.begin(), in.storage_keys.begin() + n_pairs));
    std::vector<MerkleProofInput> items;
    std::vector<int32_t> rfp;
    items.push_back({in.account_proof, in.root_hash, Bytes(in.address_keccak.begin(), in.address_keccak.end())});
    rfp.push_back(-1);
    for (size_t i = 0; i < n_pairs; i++) {
      items.push_back({in.storage_proofs[i], Bytes(32, 0), Bytes(hashed[i].begin(), hashed[i].end())});
      rfp.push_back(0);
    }
    std::vector<Outcome> res = verify_merkle_proofs(items, &rfp);
    std::vector<Bytes> values;
    for (size_t i = 0; i < res.size(); i++) {
      if (!res[i].ok()) throw VerifyPanic(res[i].status);
      // decode_exact::<Account>(..).unwrap() (main.rs:15) runs even when no storage proof follows
      if (i == 0 && !mptv_account_storage_root(res[0].value.data(), (uint32_t)res[0].value.size(), nullptr))
        throw VerifyPanic(MPTV_ST_DEP_FAILED);
      if (i) values.push_back(res[i].value);
    }
    return values;
  }

  // The storage guest over a batch of inputs that exist as bytes: blob i = blobs[off[i] .. off[i+1]) holds
  // borsh(StorageProofInput) (types.rs:11-19, main.rs:6-9).  Per input: the committed storage values, or the status
  // the guest dies with (first failing proof in its order).  One streamed call (mptv_verify_storage_borsh).
  struct StorageOutcome {
    int status = MPTV_ST_OK;
    std::vector<Bytes> values;
    bool ok() const { return status == MPTV_ST_OK; }
  };
  std::vector<StorageOutcome> verify_storage_borsh(const uint8_t* blobs, const uint64_t* off, size_t n, size_t n_proofs = 0) {
    std::vector<StorageOutcome> out(n);
    if (n == 0) return out;
    std::vector<uint64_t> first(n + 1);
    std::vector<uint8_t> ist(n), status;
    std::vector<uint64_t> voff;
    std::vector<uint32_t> vlen;
    for (size_t cap = n_proofs;;) {
      status.assign(cap, 0); voff.assign(cap, 0); vlen.assign(cap, 0);
      mptv_result r{status.data(), voff.data(), vlen.data()};
      const int rc = mptv_verify_storage_borsh(ctx_, blobs, off, n, 0, first.data(), ist.data(), cap, &r);
      if (rc == MPTV_ERR_NOMEM && first[n] > cap) { cap = (size_t)first[n]; continue; }  // proof_first holds the layout
      check(rc, "mptv_verify_storage_borsh");
      break;
    }
    for (size_t i = 0; i < n; i++) {
      out[i].status = ist[i];
      if (!out[i].ok()) continue;
      for (uint64_t p = first[i] + 1; p < first[i + 1]; p++) out[i].values.emplace_back(blobs + voff[p], blobs + voff[p] + vlen[p]);
    }
    return out;
  }
  std::vector<StorageOutcome> verify_storage_proof_inputs(const std::vector<StorageProofInput>& inputs) {
    Bytes blobs;
    std::vector<uint64_t> off{0};
    size_t n_proofs = 0;
    for (const StorageProofInput& in : inputs) {
      const Bytes b = in.to_borsh();
      blobs.insert(blobs.end(), b.begin(), b.end());
      off.push_back(blobs.size());
      n_proofs += 1 + std::min(in.storage_proofs.size(), in.storage_keys.size());
    }
    blobs.resize(blobs.size() + 16, 0);
    return verify_storage_borsh(blobs.data(), off.data(), inputs.size(), n_proofs);
  }

  void check(int rc, const char* what) {
    if (rc != MPTV_OK) throw MptvError(std::string(what) + ": " + mptv_strerror(rc) + " (" + mptv_last_error(ctx_) + ")");
  }

 private:
  mptv_ctx* ctx_ = nullptr;
  mptv_host_batch* pinned_ = nullptr;    // recycled flattener output buffers
  mptv_host_batch* pageable_ = nullptr;
};

inline Verifier& default_verifier() {
  static thread_local Verifier v(std::vector<int>{0});
  return v;
}
inline Bytes verify_merkle_proof(const B256& root_hash, std::vector<Bytes> proof, const Bytes& key) {
  return default_verifier().verify_merkle_proof(root_hash, std::move(proof), key);
}
inline std::vector<Outcome> verify_merkle_proofs(const std::vector<MerkleProofInput>& inputs) {
  return default_verifier().verify_merkle_proofs(inputs);
}
inline B256 digest_keccak(const Bytes& b) { return default_verifier().digest_keccak(b); }

}  // namespace crypto_ops

namespace trie_utils {

using crypto_ops::B256;
using crypto_ops::Bytes;

// trie-utils/src/types.rs:11-15
struct Log {
  std::array<uint8_t, 20> address{};
  std::vector<B256> topics;
  Bytes data;
};

// alloy_rlp::encode(index) (transaction.rs:45)
inline Bytes rlp_index(uint64_t index) {
  uint8_t b[9];
  uint32_t n = mptv_rlp_index(index, b);
  return Bytes(b, b + n);
}

// insert_receipt's leaf bytes (receipt.rs:8-38)
inline Bytes encode_receipt(std::optional<uint8_t> prefix, bool status, uint64_t cumulative_gas_used,
                            const std::array<uint8_t, 256>& bloom, const std::vector<Log>& logs) {
  std::vector<Bytes> topics(logs.size());
  std::vector<mptv_log> cl(logs.size());
  for (size_t i = 0; i < logs.size(); i++) {
    for (const B256& t : logs[i].topics) topics[i].insert(topics[i].end(), t.begin(), t.end());
    cl[i] = mptv_log{logs[i].address.data(), topics[i].data(), (uint32_t)logs[i].topics.size(), logs[i].data.data(),
                     (uint32_t)logs[i].data.size()};
  }
  const int p = prefix ? (int)*prefix : -1;
  uint64_t n = mptv_encode_receipt(p, status, cumulative_gas_used, bloom.data(), cl.data(), (uint32_t)cl.size(), nullptr, 0);
  Bytes out(n);
  mptv_encode_receipt(p, status, cumulative_gas_used, bloom.data(), cl.data(), (uint32_t)cl.size(), out.data(), n);
  return out;
}

namespace detail {
struct Kv {
  Bytes key_bytes, value_bytes;
  std::vector<uint32_t> key_off{0}, value_len, trie_first{0};
  std::vector<uint64_t> value_off;
  void push(const Bytes& k, const Bytes& v) {
    key_bytes.insert(key_bytes.end(), k.begin(), k.end());
    key_off.push_back((uint32_t)key_bytes.size());
    value_off.push_back(value_bytes.size());
    value_len.push_back((uint32_t)v.size());
    value_bytes.insert(value_bytes.end(), v.begin(), v.end());
    value_bytes.resize((value_bytes.size() + 15) & ~(size_t)15, 0);
  }
  void end_trie() { trie_first.push_back((uint32_t)value_len.size()); }
  mptv_kv_batch view() {
    value_bytes.resize(value_bytes.size() + 16, 0);
    key_bytes.resize(key_bytes.size() + 16, 0);
    return mptv_kv_batch{key_bytes.data(), key_off.data(), value_bytes.data(), value_bytes.size(), value_off.data(),
                         value_len.data(), value_len.size(), trie_first.data(), trie_first.size() - 1};
  }
};
}  // namespace detail

// root of {rlp(i): items[i]} -- EthTrie::new + insert x n + root_hash (transaction.rs:41-66)
inline B256 ordered_trie_root(crypto_ops::Verifier& v, const std::vector<Bytes>& items) {
  detail::Kv kv;
  for (size_t i = 0; i < items.size(); i++) kv.push(rlp_index(i), items[i]);
  kv.end_trie();
  mptv_kv_batch b = kv.view();
  B256 root{};
  v.check(mptv_trie_roots(v.ctx(), &b, root.data()), "mptv_trie_roots");
  return root;
}

// the RPC-free half of get_ethereum_transaction_proof_inputs / get_ethereum_receipt_proof_inputs
// (transaction.rs:41-73, receipt.rs:49-92): build the trie, root_hash(), get_proof(rlp(target_index))
inline crypto_ops::MerkleProofInput transaction_proof_inputs(crypto_ops::Verifier& v, const std::vector<Bytes>& items,
                                                            uint32_t target_index) {
  detail::Kv kv;
  for (size_t i = 0; i < items.size(); i++) kv.push(rlp_index(i), items[i]);
  kv.end_trie();
  mptv_kv_batch b = kv.view();
  Bytes key = rlp_index(target_index);
  uint32_t trie = 0, koff[2] = {0, (uint32_t)key.size()};
  mptv_proof_targets tg{&trie, key.data(), koff, 1};
  B256 root{};
  uint32_t proof_first[2] = {0, 0};
  std::vector<uint8_t> nb(1 << 16);
  std::vector<uint64_t> noff(64);
  std::vector<uint32_t> nlen(64);
  mptv_proofs_out out{nb.data(), nb.size(), noff.data(), nlen.data(), noff.size(), proof_first, 0, 0};
  int rc = mptv_trie_proofs(v.ctx(), &b, &tg, root.data(), &out);
  if (rc == MPTV_ERR_NOMEM) {
    nb.resize(out.node_bytes_len + 16); noff.resize(out.n_nodes + 1); nlen.resize(out.n_nodes + 1);
    out = mptv_proofs_out{nb.data(), nb.size(), noff.data(), nlen.data(), noff.size(), proof_first, 0, 0};
    rc = mptv_trie_proofs(v.ctx(), &b, &tg, root.data(), &out);
  }
  v.check(rc, "mptv_trie_proofs");
  crypto_ops::MerkleProofInput m;
  for (uint64_t i = 0; i < out.n_nodes; i++) m.proof.emplace_back(nb.begin() + noff[i], nb.begin() + noff[i] + nlen[i]);
  m.root_hash.assign(root.begin(), root.end());
  m.key = key;
  return m;
}

}  // namespace trie_utils
