// C++ tests of the reference-shaped host API (include/mptv_crypto_ops.hpp), written the way the
// reference's own tests read:
//   /root/reference/trie-utils/tests/rlp.rs:12-47          test_encode_receipt (known answer)
//   /root/reference/trie-utils/tests/transaction.rs:10-36  build proof inputs, verify_merkle_proof, compare
//   /root/reference/trie-utils/tests/storage.rs            account + storage verification
// `--cpu`: host-side codecs only (no GPU needed).  `--gpu`: everything, on cuda:0.
// Known answers for the GPU part are SURVEY.md Appendix E (outputs of the reference's own guest ELF).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mptv_crypto_ops.hpp"

using crypto_ops::B256;
using crypto_ops::Bytes;

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); g_fail++; } } while (0)

static Bytes unhex(const std::string& s) {
  Bytes b;
  for (size_t i = 0; i + 1 < s.size(); i += 2) b.push_back((uint8_t)strtol(s.substr(i, 2).c_str(), nullptr, 16));
  return b;
}
static B256 b256(const std::string& s) { B256 r{}; Bytes b = unhex(s); memcpy(r.data(), b.data(), 32); return r; }
static std::string hex(const uint8_t* p, size_t n) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; }
  return s;
}

// trie-utils/tests/rlp.rs:12
static void test_encode_receipt() {
  std::string expected = "f901668001b90100" + std::string(512, '0') +
      "f85ff85d940000000000000000000000000000000000000011f842a0"
      "000000000000000000000000000000000000000000000000000000000000deada0"
      "000000000000000000000000000000000000000000000000000000000000beef830100ff";
  trie_utils::Log log;
  Bytes a = unhex("0000000000000000000000000000000000000011");
  memcpy(log.address.data(), a.data(), 20);
  log.topics = {b256("000000000000000000000000000000000000000000000000000000000000dead"),
                b256("000000000000000000000000000000000000000000000000000000000000beef")};
  log.data = unhex("0100ff");
  std::array<uint8_t, 256> bloom{};
  Bytes out = trie_utils::encode_receipt(std::nullopt, false, 0x1, bloom, {log});
  CHECK(hex(out.data(), out.size()) == expected);
  Bytes typed = trie_utils::encode_receipt((uint8_t)0x02, false, 0x1, bloom, {log});
  CHECK(typed.size() == out.size() + 1 && typed[0] == 0x02 && Bytes(typed.begin() + 1, typed.end()) == out);
}

static void test_borsh_and_keys() {
  crypto_ops::MerkleProofInput in{{unhex("0102"), {}, Bytes(40, 7)}, Bytes(32, 0xaa), unhex("1234")};
  Bytes w = in.to_borsh();
  CHECK(crypto_ops::MerkleProofInput::from_borsh(w.data(), w.size()) == in);
  bool threw = false;
  try { crypto_ops::MerkleProofInput::from_borsh(w.data(), w.size() - 1); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  CHECK(trie_utils::rlp_index(0) == unhex("80"));
  CHECK(trie_utils::rlp_index(15) == unhex("0f"));
  CHECK(trie_utils::rlp_index(128) == unhex("8180"));
  CHECK(trie_utils::rlp_index(300) == unhex("82012c"));
  // the flattener on a small batch
  Bytes blobs = w;
  uint64_t off[2] = {0, w.size()};
  mptv_host_batch* hb = nullptr;
  CHECK(mptv_flatten_borsh(blobs.data(), off, 1, 2, 0, &hb) == MPTV_OK);
  const mptv_batch* v = mptv_host_batch_view(hb);
  CHECK(v->n_proofs == 1 && v->n_nodes == 3 && v->node_len[0] == 2 && v->node_len[1] == 0 && v->node_len[2] == 40);
  CHECK(v->node_off[0] == 0 && v->node_off[1] == 16 && v->node_off[2] == 16 && v->key_off[1] == 2);
  mptv_host_batch_free(hb);
}

// crypto-ops/src/types.rs:11-19: the borsh form of the storage guest's input, and what its host flattener makes of it
static void test_storage_input_borsh() {
  crypto_ops::StorageProofInput s;
  s.account_proof = {unhex("c482208080"), unhex("")};
  s.storage_proofs = {{unhex("c582208081ff")}, {}, {unhex("01"), unhex("0203")}};
  s.root_hash = Bytes(32, 7);
  s.account_key = unhex("aabb");
  s.storage_keys = {unhex("01"), Bytes(32, 9)};  // one key fewer than proofs: the guest's zip drops the third proof
  s.address_keccak.fill(5);
  const Bytes w = s.to_borsh();
  CHECK(crypto_ops::StorageProofInput::from_borsh(w.data(), w.size()) == s);
  bool threw = false;
  try { crypto_ops::StorageProofInput::from_borsh(w.data(), w.size() - 1); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  uint64_t off[2] = {0, w.size()}, first[2] = {9, 9};
  mptv_host_batch* hb = nullptr;
  const uint8_t* hk = nullptr;
  CHECK(mptv_flatten_storage_borsh(w.data(), off, 1, 1, 0, 0, &hb, nullptr, first, &hk) == MPTV_OK);
  const mptv_batch* v = mptv_host_batch_view(hb);
  CHECK(first[0] == 0 && first[1] == 3 && v->n_proofs == 3 && v->n_nodes == 3);  // account proof (2 nodes) + 2 storage proofs (1 + 0 nodes)
  CHECK(v->proof_first[1] == 2 && v->proof_first[2] == 3 && v->proof_first[3] == 3);
  CHECK(v->root_from_proof[0] == -1 && v->root_from_proof[1] == 0 && v->root_from_proof[2] == 0 && hk[0] == 0 && hk[1] == 1 && hk[2] == 1);
  CHECK(v->key_off[1] - v->key_off[0] == 32 && v->key_off[2] - v->key_off[1] == 1 && v->key_off[3] - v->key_off[2] == 32);
  CHECK(v->roots[0] == 7 && v->roots[32] == 0);
  mptv_host_batch_free(hb);
  // the guest's Account decode on the host
  uint8_t root[32];
  const Bytes acct = unhex("f84605820100a01111111111111111111111111111111111111111111111111111111111111111a02222222222222222222222222222222222222222222222222222222222222222");
  CHECK(mptv_account_storage_root(acct.data(), (uint32_t)acct.size(), root) == 1 && root[0] == 0x11 && root[31] == 0x11);
  CHECK(mptv_account_storage_root(acct.data(), (uint32_t)acct.size() - 1, nullptr) == 0);
}

static void test_no_gpu_is_loud() {
  bool threw = false;
  try { crypto_ops::Verifier v({0}); } catch (const crypto_ops::MptvError&) { threw = true; }
  CHECK(threw);
}

// SURVEY.md Appendix E KAT-1,2,3,4,6,8,11,12,14 -- produced by the reference ELF
static void test_verify_merkle_proof_known_answers() {
  using crypto_ops::verify_merkle_proof;
  const B256 r1 = b256("0e9985286c0f4a35519eeb86fa50ce8134ed1fd1bb88e6f74a9e4cb6f505079c");
  CHECK(verify_merkle_proof(r1, {unhex("cc822080880102030405060708")}, unhex("80")) == unhex("0102030405060708"));
  auto panics_with = [](int status, const char* msg, auto fn) {
    try { fn(); } catch (const crypto_ops::VerifyPanic& e) { return e.status == status && std::string(e.what()).find(msg) != std::string::npos; }
    return false;
  };
  CHECK(panics_with(MPTV_ST_KEY_NOT_FOUND, "Key does not exist!", [&] { verify_merkle_proof(r1, {unhex("cc822080880102030405060708")}, unhex("01")); }));
  CHECK(panics_with(MPTV_ST_ROOT_NOT_CANONICAL, "left == right", [&] {
    verify_merkle_proof(b256("597b7d6dac7ed0e172717eb3ecee0ebd56188817afe42be254e47a5d322efc89"), {unhex("c582208081ff")}, unhex("80")); }));
  CHECK(verify_merkle_proof(b256("1e03594df303045ca22e7d8f7ff4504ea2a9e1168864bf7b16ec37b79d9e2671"), {unhex("c482208080")}, unhex("80")).empty());
  const B256 r6 = b256("3888c5866c792987e82c5b40231486a304f99b1fce762a1f5a6fd9e798f990ba");
  const Bytes n6 = unhex("d780c43082aabbc230058080808080808080808080808080");
  CHECK(verify_merkle_proof(r6, {n6}, unhex("10")) == unhex("aabb"));
  CHECK(panics_with(MPTV_ST_KEY_NOT_FOUND, "Key does not exist!", [&] { verify_merkle_proof(r6, {n6}, unhex("30")); }));
  const B256 r11 = b256("48245dc08200ffbefb3f22a59d0b178b5f4bb15fe07545be92e30723098841b5");
  const Bytes leaf = unhex("e43ca2404142434445464748494a4b4c4d4e4f505152535455565758595a5b5c5d5e5f6061");
  const Bytes branch = unhex("f180808080808080808080a03f68b9461a405e7e99ba7734a1836f01ac1f95840617f8f8d6bcde48dbb0f9ef808080808080");
  CHECK(verify_merkle_proof(r11, {leaf, unhex("c0"), branch}, unhex("ac")) == unhex("404142434445464748494a4b4c4d4e4f505152535455565758595a5b5c5d5e5f6061"));
  CHECK(panics_with(MPTV_ST_INVALID_PROOF, "InvalidProof", [&] { verify_merkle_proof(r11, {branch}, unhex("ac")); }));
  B256 r14 = r11; r14[0] = 0x49;
  CHECK(panics_with(MPTV_ST_INVALID_STATE_ROOT, "InvalidStateRoot", [&] { verify_merkle_proof(r14, {branch, leaf}, unhex("ac")); }));
  CHECK(hex(crypto_ops::digest_keccak(Bytes{}).data(), 32) == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470");
}

// trie-utils/tests/transaction.rs: build the proof inputs for one index, verify, compare with the inserted bytes
static void test_transaction_proof_roundtrip() {
  crypto_ops::Verifier& v = crypto_ops::default_verifier();
  std::vector<Bytes> txs;
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < 200; i++) {
    Bytes t(100 + (i * 37) % 200, 0);
    for (auto& b : t) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b = (uint8_t)s; }
    t[0] = 0x02;
    txs.push_back(t);
  }
  CHECK(hex(trie_utils::ordered_trie_root(v, {unhex("0102030405060708")}).data(), 32) ==
        "0e9985286c0f4a35519eeb86fa50ce8134ed1fd1bb88e6f74a9e4cb6f505079c");  // KAT-1's trie
  for (uint32_t target : {15u, 0u, 127u, 128u, 199u}) {
    crypto_ops::MerkleProofInput in = trie_utils::transaction_proof_inputs(v, txs, target);
    B256 root{};
    memcpy(root.data(), in.root_hash.data(), 32);
    CHECK(root == trie_utils::ordered_trie_root(v, txs));
    CHECK(crypto_ops::verify_merkle_proof(root, in.proof, in.key) == txs[target]);
  }
  crypto_ops::MerkleProofInput absent = trie_utils::transaction_proof_inputs(v, txs, 200);
  std::vector<crypto_ops::Outcome> o = crypto_ops::verify_merkle_proofs({absent});
  CHECK(o.size() == 1 && o[0].status == MPTV_ST_KEY_NOT_FOUND);
}

// the batched entry on a batch large enough for the streamed borsh path (> 8 MB of blobs): every tx of 40 blocks
// proven and verified, plus one absent index per block, against the element-wise answers
// the storage guest's batched mirror through the wire format: inputs whose "account" leaf is a transaction (no
// Account), a wrong root, a 31-byte root -- what the guest dies with, input by input
static void test_storage_inputs_through_the_wire_format() {
  crypto_ops::Verifier& v = crypto_ops::default_verifier();
  std::vector<Bytes> txs;
  uint64_t s = 424242;
  for (int i = 0; i < 60; i++) {
    Bytes t(120 + (i * 11) % 90, 0);
    for (auto& b : t) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b = (uint8_t)s; }
    t[0] = 0x02;
    txs.push_back(t);
  }
  std::vector<crypto_ops::StorageProofInput> ins;
  for (uint32_t target : {3u, 17u, 59u, 60u}) {
    crypto_ops::MerkleProofInput m = trie_utils::transaction_proof_inputs(v, txs, target);
    crypto_ops::StorageProofInput in;
    in.account_proof = m.proof;
    in.root_hash = m.root_hash;
    // (a transaction trie is keyed by rlp(index), not by a 32-byte hash: under address_keccak the key is simply absent)
    in.address_keccak.fill(0);
    memcpy(in.address_keccak.data(), m.key.data(), std::min<size_t>(32, m.key.size()));
    in.storage_proofs = {m.proof};
    in.storage_keys = {unhex("01")};
    ins.push_back(in);
  }
  ins[1].root_hash[0] ^= 1;
  ins[2].root_hash.pop_back();
  std::vector<crypto_ops::Verifier::StorageOutcome> o = v.verify_storage_proof_inputs(ins);
  CHECK(o.size() == 4);
  for (size_t i = 0; i < o.size(); i++) {
    int want = -1;
    try { v.verify_storage_proof_input(ins[i]); want = MPTV_ST_OK; } catch (const crypto_ops::VerifyPanic& e) { want = e.status; }
    CHECK(o[i].status == want && want != MPTV_ST_OK);
  }
  CHECK(o[1].status == MPTV_ST_INVALID_STATE_ROOT && o[2].status == MPTV_ST_BAD_ROOT_LEN);
}

static void test_large_batch_takes_the_streamed_path() {
  crypto_ops::Verifier& v = crypto_ops::default_verifier();
  std::vector<crypto_ops::MerkleProofInput> inputs;
  std::vector<Bytes> want;
  uint64_t s = 1234567;
  for (int blk = 0; blk < 40; blk++) {
    std::vector<Bytes> txs;
    for (int i = 0; i < 150; i++) {
      Bytes t(300 + (i * 53 + blk * 7) % 900, 0);
      for (auto& b : t) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b = (uint8_t)s; }
      t[0] = 0x02;
      txs.push_back(t);
    }
    for (uint32_t target = 0; target <= 150; target++) {  // 150 is absent
      inputs.push_back(trie_utils::transaction_proof_inputs(v, txs, target));
      want.push_back(target < 150 ? txs[target] : Bytes{});
    }
  }
  size_t blob_bytes = 0;
  for (auto& in : inputs) blob_bytes += in.to_borsh().size();
  CHECK(blob_bytes > (8u << 20));
  std::vector<crypto_ops::Outcome> o = v.verify_merkle_proofs(inputs);
  CHECK(o.size() == inputs.size());
  for (size_t i = 0; i < o.size(); i++) {
    if (want[i].empty()) CHECK(o[i].status == MPTV_ST_KEY_NOT_FOUND);
    else CHECK(o[i].ok() && o[i].value == want[i]);
  }
  inputs[5].root_hash.pop_back();  // 31-byte root: the guests' try_into().unwrap()
  o = v.verify_merkle_proofs(inputs);
  CHECK(o[5].status == MPTV_ST_BAD_ROOT_LEN && o[4].ok() && o[6].ok());
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "--cpu";
  test_encode_receipt();
  test_borsh_and_keys();
  test_storage_input_borsh();
  if (mode == "--cpu-nogpu") test_no_gpu_is_loud();
  if (mode == "--gpu") {
    test_verify_merkle_proof_known_answers();
    test_transaction_proof_roundtrip();
    test_large_batch_takes_the_streamed_path();
    test_storage_inputs_through_the_wire_format();
  }
  printf("%s: %d failure(s)\n", mode.c_str(), g_fail);
  return g_fail ? 1 : 0;
}
