// kernels.h -- internal launch interface between the CUDA kernels and the C-ABI layer (mptv_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mptv {

constexpr uint32_t kMaxNodeLen = 0xffff0000u;  // 32-bit byte positions inside a node (136 k + 136) must not wrap
constexpr int kNumBins = 128;            // rate-block-count bins (K0)
constexpr int kLongLeafBin = 33;         // nodes of more than 32 rate blocks (> 4.3 KB) are hashed in a launch of their own
constexpr int kBinScratchWords = 2 * kNumBins + 4;  // hist | cursor | K1 tile counter
constexpr int kBinNodesPerBlock = 4096;  // nodes handled by one CTA of the binning kernels
constexpr int kWalkThreads = 256;        // K2b CTA size (the latency kernel runs the same walk with fewer threads)
constexpr int kKeccakThreads = 128;      // K1 CTA size: one node per thread
constexpr int kKeccakMinBlocks = 4;      // resident CTAs / SM  (=> <= 128 registers / thread)

// K2a keeps this many DFS frames (a window over the innermost levels; deeper nests are re-walked from the top of
// the node when the decode returns past the window -- verify_kernels.cu replay_frames).  Power of two.
constexpr int kInlineWindow = 64;

// verdict classes, 1:1 with the reference's outcomes (include/mptv.h MPTV_ST_*)
enum : uint32_t {
  kStOk = 0, kStInvalidStateRoot = 1, kStRootNotCanonical = 2, kStInvalidProof = 3,
  kStKeyNotFound = 4, kStPanicOther = 5, kStBadRootLen = 6, kStDepFailed = 7
};
// K2a per-node record: kind[0:3) dec[3:5) canon[5] fast[6] hdr_len[7:10) mask16[10:26)
enum : uint32_t { kKindEmpty = 0, kKindLeaf = 1, kKindExt = 2, kKindBranch = 3, kKindHash = 4 };
enum : uint32_t { kDecOk = 0, kDecErr = 1, kDecPanic = 2 };
__host__ __device__ __forceinline__ uint32_t make_meta(uint32_t kind, uint32_t dec, uint32_t canon, uint32_t fast,
                                                       uint32_t hdr_len, uint32_t mask) {
  return kind | (dec << 3) | (canon << 5) | (fast << 6) | (hdr_len << 7) | (mask << 10);
}
constexpr uint32_t kMetaSlow = 0xffffffffu;  // K1 could not classify the node: k_parse_nodes decides
__host__ __device__ __forceinline__ uint32_t meta_kind(uint32_t m) { return m & 7u; }
__host__ __device__ __forceinline__ uint32_t meta_dec(uint32_t m) { return (m >> 3) & 3u; }
__host__ __device__ __forceinline__ uint32_t meta_canon(uint32_t m) { return (m >> 5) & 1u; }
__host__ __device__ __forceinline__ uint32_t meta_fast(uint32_t m) { return (m >> 6) & 1u; }
__host__ __device__ __forceinline__ uint32_t meta_hdr(uint32_t m) { return (m >> 7) & 7u; }
__host__ __device__ __forceinline__ uint32_t meta_mask(uint32_t m) { return (m >> 10) & 0xffffu; }

// device-resident batch in the CSR layout of include/mptv.h
struct DeviceBatch {
  const uint8_t* node_bytes;
  const uint64_t* node_off;
  const uint32_t* node_len;
  uint64_t n_nodes;
  const uint32_t* proof_first;
  uint64_t n_proofs;
  const uint8_t* roots;
  const uint8_t* key_bytes;
  const uint32_t* key_off;
  const uint32_t* key_len;         // null: key p = [key_off[p], key_off[p+1]); else key_len[p] bytes at key_off[p] (keys placed freely)
  const int32_t* root_from_proof;  // may be null
  // slice view of a larger CSR (host-buffer pipeline): the arrays above hold the slice, but their
  // VALUES are still global indices / offsets; these bases translate them.  All 0 for a whole batch.
  uint64_t byte_base;   // node_off values are relative to node_bytes - byte_base
  uint32_t node_base;   // proof_first values are node indices + node_base
  uint32_t key_base;    // key_off values are offsets + key_base
  uint64_t proof_base;  // root_from_proof values are proof indices + proof_base
};

// per-device one-time setup (dynamic shared memory opt-in); call with the device current
cudaError_t kernels_init_device();

// K0: order[] = node indices sorted by DESCENDING rate-block bin.  scratch = 2*kNumBins u32.
// ids == NULL: nodes 0..n_nodes-1; else the n_nodes node indices listed in ids[].
// keep != NULL: only nodes with keep[i] == i are binned (order[] then holds just those) and totals[0..1]
// (device) receive their count and Keccak-f count.
cudaError_t launch_bin_nodes(const uint32_t* node_len, const uint32_t* ids, uint64_t n_nodes, uint32_t* scratch,
                             uint32_t* order, cudaStream_t st, const uint32_t* keep = nullptr,
                             unsigned long long* totals = nullptr, int long_bin = kLongLeafBin);
// the device word (inside launch_bin_nodes' scratch) that holds the number of nodes in bins >= long_bin
inline const uint32_t* bin_split_word(const uint32_t* scratch) { return scratch + 2 * kNumBins + 1; }
// K1: digests[32*i] = keccak256(node i).  order may be NULL (identity).  meta may be NULL; when
// given, meta[i] receives the K2a record of plain branches / plain leaves and kMetaSlow otherwise.
cudaError_t launch_keccak256_nodes(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off,
                                   const uint32_t* node_len, const uint32_t* order, uint64_t n_nodes,
                                   uint8_t* digests, uint32_t* meta, uint32_t* tile_counter /* 1 u32 of scratch or NULL */,
                                   int sm_count, cudaStream_t st, const uint32_t* split = nullptr, int long_ctas = 1);

// storage guest: per-proof key records (offset from key_bytes, length); the keys of proofs with hash_key[p] != 0 are
// replaced by their Keccak-256, written to hashed[32 p ..] (hashed_off = hashed - key_bytes, same allocation)
cudaError_t launch_prepare_keys(const uint8_t* key_bytes, const uint32_t* key_off, uint32_t key_base, const uint8_t* hash_key,
                                uint64_t n_proofs, uint8_t* hashed, uint32_t hashed_off, uint32_t* off_out, uint32_t* len_out,
                                cudaStream_t st, const uint32_t* key_len_in = nullptr /* key_off holds (offset) records, this the lengths */);

// pull mode of the streamed borsh entry: records (src offset u64, dst / 16 u32, len u32; len 0xffffffff = unused) ->
// dst_base + 16 dst16 receives len bytes of src_base + src, zero-padded to a multiple of 16.  src_base is mapped
// page-locked HOST memory (the caller's blobs): the loads cross PCIe.
cudaError_t launch_gather(const uint8_t* src_base, uint8_t* dst_base, const uint4* recs, uint32_t n_recs, int sm_count,
                          cudaStream_t st);

// device flatten (borsh_kernels.cu): a chunk of borsh(MerkleProofInput) blobs, resident on the device as the caller
// wrote them (img, blob i = [off[i], off[i + 1])), becomes the CSR arrays of a DeviceBatch.
// count: per-blob node count / padded bytes / flags (bit 0 well-formed, bit 1 root_hash.len() != 32), their exclusive
// scans, and totals = {nodes, bytes, malformed blobs}.  emit (only when no blob is malformed): index arrays and one
// gather record per node and key for launch_gather.  map: results in place -> offsets into the caller's blobs.
cudaError_t launch_blob_count(const uint8_t* img, const uint64_t* off, uint32_t np, uint32_t* n_nodes, uint64_t* n_bytes, uint8_t* flags,
                              uint32_t* node_first, uint64_t* byte_first, unsigned long long* totals, cudaStream_t st);
cudaError_t launch_blob_emit(const uint8_t* img, const uint64_t* off, uint32_t np, uint32_t nn, const uint32_t* node_first,
                             const uint64_t* byte_first, uint64_t arena_off, uint64_t blob_base, uint64_t* node_off, uint32_t* node_len,
                             uint64_t* node_src, uint32_t* proof_first, uint8_t* roots, uint32_t* key_off, uint32_t* key_len,
                             uint4* recs, cudaStream_t st);
cudaError_t launch_blob_map(uint32_t np, const uint8_t* flags, const uint32_t* proof_first, const uint64_t* node_off,
                            const uint32_t* node_len, const uint64_t* node_src, uint8_t* status, uint64_t* value_off,
                            uint32_t* value_len, cudaStream_t st);

// K2a: meta[i] = eager-decode record of node i.  only_slow: leave records != kMetaSlow untouched
cudaError_t launch_parse_nodes(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off,
                               const uint32_t* node_len, uint64_t n_nodes, uint32_t* meta, bool only_slow,
                               cudaStream_t st);
// K2f + K2b: the thread-per-proof chain check decides what it can, K2b the deferred rest.
// wave 0 = proofs with their own root, wave 1 = proofs whose root is another proof's storage_root
cudaError_t launch_verify_walk(const DeviceBatch& b, const uint8_t* digests, const uint32_t* meta, int wave,
                               int lanes_per_proof, uint8_t* status, uint64_t* value_off, uint32_t* value_len,
                               uint32_t* defer /* scratch [1 + n_proofs]; NULL = K2b on every proof */, int sm_count,
                               cudaStream_t st);

// ------------------------------------------------------------------ latency path (single_kernels.cu)
// A batch of at most kSmallMaxNodes nodes / kSmallMaxProofs proofs whose packed form fits kSmallMaxPack bytes is
// verified by ONE launch of one CTA, inputs and results through mapped page-locked memory.
constexpr int kSmallThreads = 128;       // one warp per scheduler: the critical path is one thread's Keccak chain
constexpr uint32_t kSmallMaxNodes = 128, kSmallMaxProofs = 32;
constexpr uint32_t kSmallMaxPack = 64 << 10;
constexpr int kSmallMaxSmem = 96 << 10;  // pack + guard + digests + records + results
constexpr uint32_t kSmallOutBytes = 16 + 13 * kSmallMaxProofs + 16 + 64;  // + room for the optional phase clocks (MPTV_SMALL_TIMING)
struct SmallHeader {
  uint32_t n_nodes, n_proofs, total;                             // total = bytes of the pack
  uint32_t o_bytes, o_off, o_len, o_pf, o_roots, o_keys, o_koff, o_rfp, has_rfp;  // arrays inside the pack
  uint32_t o_hk, has_hk, hk_scratch;                             // hash_key flags (hashed-keys entry) and the kernel's key records
  uint32_t scratch, results;                                     // shared-memory offsets of the kernel's own arrays
  uint32_t seq;                                                  // written to the mailbox when the results are there
  uint32_t node_base, key_base;
  uint64_t byte_base, proof_base;                                // the slice translation of DeviceBatch
};
cudaError_t small_init_device();
cudaError_t launch_verify_small(const uint8_t* mailbox_dev, const SmallHeader& h, uint8_t* out_dev, int lanes_per_proof,
                                cudaStream_t st);

// ------------------------------------------------------------------ K4: trie rebuild (rebuild_kernels.cu)
constexpr int kTrieThreads = 128;       // CTA of k_trie_structure (one trie per CTA)
constexpr int kTrieMaxItems = 8192;     // items per trie (shared-memory sort); larger tries are refused
constexpr int kTrieMaxKeyLen = 32;      // key bytes (every Ethereum trie key is <= 32 bytes)
constexpr int kMaxLevels = 136;         // node heights: < 2 * 64 nibbles + 2

// device-resident key/value batch in the layout of include/mptv.h `mptv_kv_batch`
struct TrieBatchDev {
  const uint8_t* key_bytes;
  const uint32_t* key_off;    // [n_items + 1]
  const uint8_t* value_bytes;
  const uint64_t* value_off;  // [n_items], 16-byte aligned
  const uint32_t* value_len;  // [n_items], 0 = delete
  const uint32_t* trie_first; // [n_tries + 1]
  uint32_t n_tries;
  uint64_t n_items;
};
struct TrieSummary {
  unsigned long long arena_bytes, perms, nodes_hashed, n_nodes;
  uint32_t max_items, max_key_len, error, pad;
  uint32_t lvl_count[2 * kMaxLevels];   // [2h] hashed nodes of height h, [2h+1] inline (< 32 byte) nodes
  uint32_t lvl_cursor[2 * kMaxLevels];
};
// scratch; node id = 3 * trie_first[t] + BFS index (a trie of n items has < 3n nodes)
struct TrieWork {
  uint4* rec;         // [3N] node records
  uint64_t* off;      // [3N] arena offset of the node's encoding
  uint32_t* len;      // [3N] encoded length
  uint8_t* digests;   // [32 * 3N]
  uint32_t* tcount;   // [T] nodes of trie t
  uint32_t* lvl_list; // [3N] node ids grouped by (height, hashed?)
  TrieSummary* sum;
};
cudaError_t trie_init_device();
// K1L: digests of the level-0 hashed leaves straight from the key/value arrays (no materialised encoding)
cudaError_t launch_keccak256_leaves(const TrieBatchDev& in, const uint4* rec, const uint32_t* node_len,
                                    const uint32_t* order, uint32_t n_nodes, uint8_t* digests, uint32_t* tile_counter,
                                    int sm_count, cudaStream_t st, const uint32_t* split = nullptr, int long_ctas = 1);
cudaError_t launch_trie_scan_input(const TrieBatchDev& in, TrieSummary* sum, cudaStream_t st);
cudaError_t launch_trie_structure(const TrieBatchDev& in, const TrieWork& w, uint32_t max_items, cudaStream_t st);
cudaError_t launch_trie_encode(const TrieBatchDev& in, const TrieWork& w, const uint32_t* list, uint32_t n_list,
                               uint8_t* arena, cudaStream_t st);
cudaError_t launch_trie_roots(const TrieBatchDev& in, const TrieWork& w, uint8_t* roots32, cudaStream_t st);
// get_proof after a rebuild: per-target node counts / padded bytes and their exclusive scans, then the copy
cudaError_t launch_trie_proof_count(const TrieBatchDev& in, const TrieWork& w, const uint32_t* target_trie,
                                    const uint8_t* tkey_bytes, const uint32_t* tkey_off, uint32_t n_targets, uint32_t* cnt,
                                    uint64_t* bytes, uint32_t* proof_first, uint64_t* byte_first, cudaStream_t st);
cudaError_t launch_trie_proof_emit(const TrieBatchDev& in, const TrieWork& w, const uint8_t* arena, const uint32_t* target_trie,
                                   const uint8_t* tkey_bytes, const uint32_t* tkey_off, uint32_t n_targets,
                                   const uint32_t* proof_first, const uint64_t* byte_first, uint8_t* out_bytes,
                                   uint64_t* out_off, uint32_t* out_len, bool leaves_in_arena, cudaStream_t st);

// optional node de-duplication in front of K1 (dedup_kernels.cu): dup_of[i] = the node whose digest node i
// shares (itself when unique)
cudaError_t launch_dedup_find(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off, const uint32_t* node_len,
                              uint32_t n_nodes, unsigned long long* keys, uint32_t* vals, uint32_t table_size /* power of 2 */,
                              uint32_t* slot_of, uint32_t* dup_of, cudaStream_t st);
cudaError_t launch_dedup_scatter(uint32_t n_nodes, const uint32_t* dup_of, uint8_t* digests, uint32_t* meta, cudaStream_t st);

// integer issue-rate probe (microbench.cu): mode 0 = LOP3, 1 = SHF, 2 = Keccak mix
cudaError_t run_int_peak(int mode, int sm_count, uint32_t* scratch, cudaStream_t st, double* ops_per_s);

}  // namespace mptv
