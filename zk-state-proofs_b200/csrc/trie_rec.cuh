// trie_rec.cuh -- node records of the trie rebuild (K4) and the RLP / hex-prefix size and byte rules
// shared by the structure / encode kernels (rebuild_kernels.cu) and the fused leaf hash (keccak_kernels.cu).
// Encoding rules: eth_trie write_node / Nibbles::encode_compact and alloy-rlp headers, as used by
// /root/reference/trie-utils/src/proofs/transaction.rs:41-66 through EthTrie::insert / root_hash.
#pragma once
#include <stdint.h>

namespace mptv {
namespace {

// ---- record packing: x = kind | height << 8 | path_start << 16 | path_len << 24
//                      y = item (global item index; leaf: its key/value, ext: any key below it,
//                          branch: the key whose value sits in the branch, or kNoItem)
//                      z = first child (global node id); children are consecutive in BFS order
//                      w = occupancy mask (branch) | hashed << 16
constexpr uint32_t kNoItem = 0xffffffffu;
constexpr uint32_t kPending = 0xffffffffu;  // w of a BFS queue entry that has not been expanded yet
enum : uint32_t { kTLeaf = 1, kTExt = 2, kTBranch = 3 };

__device__ __forceinline__ uint32_t rec_kind(const uint4& r) { return r.x & 0xffu; }
__device__ __forceinline__ uint32_t rec_height(const uint4& r) { return (r.x >> 8) & 0xffu; }
__device__ __forceinline__ uint32_t rec_ps(const uint4& r) { return (r.x >> 16) & 0xffu; }
__device__ __forceinline__ uint32_t rec_pl(const uint4& r) { return r.x >> 24; }
__device__ __forceinline__ uint32_t rec_mask(const uint4& r) { return r.w & 0xffffu; }
__device__ __forceinline__ uint32_t rec_hashed(const uint4& r) { return (r.w >> 16) & 1u; }

// ---- RLP sizes (alloy-rlp / eth_trie write_node)
__device__ __forceinline__ uint32_t hdr_size(uint32_t n) {
  return n < 56 ? 1u : (n < 256 ? 2u : (n < 65536 ? 3u : (n < (1u << 24) ? 4u : 5u)));
}
__device__ __forceinline__ uint32_t str_item_size(uint32_t n, uint32_t first) {
  return (n == 1 && first < 0x80) ? 1u : hdr_size(n) + n;
}
// hex-prefix encoded path of pl nibbles: pl/2 + 1 bytes (<= 33); a single byte is < 0x80 (flags <= 3)
__device__ __forceinline__ uint32_t hp_item_size(uint32_t pl) { return pl < 2 ? 1u : 2u + pl / 2; }
__device__ __forceinline__ uint32_t ref_size(uint32_t child_len) { return child_len < 32 ? child_len : 33u; }
// payload length of a list whose whole encoding is `len` bytes
__device__ __forceinline__ uint32_t payload_of(uint32_t len) {
  return len - 1 < 56 ? len - 1 : (len - 2 < 256 ? len - 2 : (len - 3 < 65536 ? len - 3 : (len - 4 < (1u << 24) ? len - 4 : len - 5)));
}
__device__ __forceinline__ uint32_t put_hdr(uint8_t* o, uint32_t n, bool list) {
  const uint32_t base = list ? 0xC0u : 0x80u;
  if (n < 56) { o[0] = (uint8_t)(base + n); return 1; }
  const uint32_t k = hdr_size(n) - 1;
  o[0] = (uint8_t)(base + 55 + k);
  for (uint32_t i = 0; i < k; i++) o[1 + i] = (uint8_t)(n >> (8 * (k - 1 - i)));
  return 1 + k;
}

// nibble i of item's key
__device__ __forceinline__ uint32_t item_nib(const uint8_t* key, uint32_t i) {
  const uint32_t b = key[i >> 1];
  return (i & 1) ? (b & 15u) : (b >> 4);
}

// byte i of the hex-prefix encoding of nibbles [ps, ps+pl) of key (Nibbles::encode_compact)
__device__ __forceinline__ uint32_t hp_byte(const uint8_t* key, uint32_t ps, uint32_t pl, bool leaf, uint32_t i) {
  const uint32_t odd = pl & 1u;
  if (i == 0) return (leaf ? 0x20u : 0u) | (odd ? (0x10u | item_nib(key, ps)) : 0u);
  const uint32_t q = ps + odd + 2 * (i - 1);
  return (item_nib(key, q) << 4) | item_nib(key, q + 1);
}


}  // namespace
}  // namespace mptv
