// verify_kernels.cu -- K2a (per-node eager RLP decode / canonical-form check), K2f (thread-per-proof
// check of chain-shaped proofs, the common case) and K2b (group-of-lanes-per-proof nibble walk with
// ballot/shuffle hash-link lookup: the full rule set, run on whatever K2f defers).
//
// Together they replace, for a whole batch, what crypto_ops::verify_merkle_proof
// (/root/reference/crypto-ops/src/lib.rs:8-23) does per proof after hashing:
//   lib.rs:14  EthTrie::from(proof_db, root)      -> root lookup by digest + eager decode_node
//   lib.rs:19  assert_eq!(root, trie.root_hash()) -> re-encode(decode(root)) == root bytes  (K3)
//   lib.rs:20  trie.verify_proof(root, key, proof)-> len>=32 filter, Nibbles::from_raw, get_at
//   lib.rs:21-22 the two expect()s                -> INVALID_PROOF / KEY_NOT_FOUND
// The third-party behaviour (eth_trie@ade617b decode_node / get_at / write_node, alloy-rlp
// Header::decode) is the rule set R1..R20 of SURVEY.md Appendix A, pinned against the
// reference's own guest ELF through oracle/ (tests/golden/verify_vectors.json).
//
// K2a runs one thread per supplied node and records, in one u32 per node, what a visit of that
// node by the reference would produce: decode status (ok / TrieError / raw panic, first failure in
// the reference's decode order), node kind, whether it survives the root-only re-encode check,
// and for plain branches the 16-bit occupancy map from which every child offset follows.
// K2b walks each proof with a group of G lanes (G = 8, 16 or 32): lane j owns the digest of node
// j; a child reference is found by comparing the 32-byte link against all digests at once
// (shuffle-broadcast of the link words + ballot), leaf / extension paths are compared
// nibble-parallel across the lanes, and the occupancy map turns the 16-way child selection into
// a popcount.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"
#include "verify_device.cuh"

namespace mptv {

__global__ void __launch_bounds__(256) k_parse_nodes(const uint8_t* __restrict__ node_bytes, uint64_t byte_base,
                                                     const uint64_t* __restrict__ node_off,
                                                     const uint32_t* __restrict__ node_len, uint64_t n_nodes,
                                                     uint32_t* __restrict__ meta, bool only_slow) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  if (only_slow && meta[i] != kMetaSlow) return;  // already decided by the fused fast path in K1
  meta[i] = parse_node(node_bytes + (node_off[i] - byte_base), node_len[i]);
}

// list == NULL: every proof of the batch (filtered by wave); else the *count proofs K2f deferred.
// Persistent: each group of G lanes strides over its share of the work list.
template <int G>
__global__ void __launch_bounds__(256, MPTV_WALK_MINB)
k_verify_walk(const DeviceBatch b, int wave, const uint8_t* __restrict__ digests,
              const uint32_t* __restrict__ meta, uint8_t* status_out, uint64_t* value_off_out,
              uint32_t* value_len_out, const uint32_t* __restrict__ list, const uint32_t* __restrict__ count) {
  const Group<G> g;
  const uint64_t n = list ? (uint64_t)*count : b.n_proofs;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x / G;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += stride)  // uniform per group
    walk_one<G>(b, wave, g, list ? (uint64_t)list[i] : i, digests, meta, status_out, value_off_out, value_len_out);
}

__global__ void __launch_bounds__(256) k_verify_fast(const DeviceBatch b, int wave, const uint8_t* __restrict__ digests,
                                                     const uint32_t* __restrict__ meta, uint8_t* status_out,
                                                     uint64_t* value_off_out, uint32_t* value_len_out,
                                                     uint32_t* __restrict__ defer_list, uint32_t* __restrict__ defer_count) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool defer = false;
  if (p < b.n_proofs) {
    const bool dependent = b.root_from_proof != nullptr && b.root_from_proof[p] >= 0;
    if (dependent == (wave == 1))
      defer = !fast_one(b, p, dependent, digests, meta, status_out, value_off_out, value_len_out);
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, defer);
  if (bal) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(bal) - 1)) base = atomicAdd(defer_count, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (defer) defer_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)p;
  }
}

// ------------------------------------------------------------------ host launchers
cudaError_t launch_parse_nodes(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off,
                               const uint32_t* node_len, uint64_t n_nodes, uint32_t* meta, bool only_slow,
                               cudaStream_t st) {
  if (n_nodes == 0) return cudaSuccess;
  unsigned blocks = (unsigned)((n_nodes + 255) / 256);
  k_parse_nodes<<<blocks, 256, 0, st>>>(node_bytes, byte_base, node_off, node_len, n_nodes, meta, only_slow);
  return cudaGetLastError();
}

cudaError_t launch_verify_walk(const DeviceBatch& b, const uint8_t* digests, const uint32_t* meta, int wave,
                               int lanes_per_proof, uint8_t* status, uint64_t* value_off, uint32_t* value_len,
                               uint32_t* defer /* [1 + n_proofs] or NULL = no fast path */, int sm_count, cudaStream_t st) {
  if (b.n_proofs == 0) return cudaSuccess;
  const int G = lanes_per_proof;
  const uint64_t threads = b.n_proofs * (uint64_t)G;
  uint64_t blocks = (threads + 255) / 256;
  const uint32_t* list = nullptr;
  const uint32_t* count = nullptr;
  if (defer) {
    cudaError_t e = cudaMemsetAsync(defer, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_verify_fast<<<(unsigned)((b.n_proofs + 255) / 256), 256, 0, st>>>(b, wave, digests, meta, status, value_off, value_len,
                                                                        defer + 1, defer);
    list = defer + 1;
    count = defer;
    blocks = std::min<uint64_t>(blocks, (uint64_t)sm_count * MPTV_WALK_MINB);  // the deferred list is short: one resident wave
  }
#define MPTV_WALK(GG) \
  k_verify_walk<GG><<<(unsigned)blocks, 256, 0, st>>>(b, wave, digests, meta, status, value_off, value_len, list, count)
  if (G == 8) MPTV_WALK(8);
  else if (G == 16) MPTV_WALK(16);
  else MPTV_WALK(32);
#undef MPTV_WALK
  return cudaGetLastError();
}

}  // namespace mptv
