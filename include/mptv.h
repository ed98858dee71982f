/*
 * mptv.h -- C ABI of the B200 batched Merkle-Patricia-Trie proof verifier.
 *
 * This is the drop-in boundary for the one hot path of ChainSafe/zk-state-proofs:
 *
 *   crypto_ops::verify_merkle_proof(root_hash: B256, proof: Vec<Vec<u8>>, key: &[u8]) -> Vec<u8>
 *       /root/reference/crypto-ops/src/lib.rs:8-23
 *   crypto_ops::keccak::digest_keccak(bytes: &[u8]) -> [u8; 32]
 *       /root/reference/crypto-ops/src/keccak.rs:6-12
 *   crypto_ops::types::{MerkleProofInput, StorageProofInput}
 *       /root/reference/crypto-ops/src/types.rs:4-19
 *   trie-utils tx / receipt trie rebuild (EthTrie::new / insert / root_hash)
 *       /root/reference/trie-utils/src/proofs/transaction.rs:41-66, proofs/receipt.rs:49-84
 *
 * The reference has no FFI of its own (it is a plain Rust function); these are the entry points a
 * Rust `extern "C"` shim binds so that `verify_merkle_proof` and the new batched
 * `verify_merkle_proofs(&[MerkleProofInput])` run on the GPU (INTEGRATION.md shows the binding).
 * Plain pointers and sizes only.  Functions return 0 or a negative MPTV_ERR_*; they never throw
 * and never abort.  One mptv_ctx is not thread-safe; separate contexts are.
 *
 * There is NO CPU fallback: without a usable CUDA device mptv_create fails.
 */
#ifndef MPTV_H
#define MPTV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-proof verdicts: 1:1 with the reference's outcomes, evaluated in the reference's order */
#define MPTV_ST_OK 0                 /* value returned                                   lib.rs:20-22 */
#define MPTV_ST_INVALID_STATE_ROOT 1 /* "Invalid merkle proof: ..."                      lib.rs:14    */
#define MPTV_ST_ROOT_NOT_CANONICAL 2 /* assert_eq!(root_hash, trie.root_hash())          lib.rs:19    */
#define MPTV_ST_INVALID_PROOF 3      /* "Failed to verify Merkle Proof: InvalidProof"    lib.rs:21    */
#define MPTV_ST_KEY_NOT_FOUND 4      /* "Key does not exist!"                            lib.rs:22    */
#define MPTV_ST_PANIC_OTHER 5        /* raw panics inside eth_trie (Nibbles::from_compact ...)        */
#define MPTV_ST_BAD_ROOT_LEN 6       /* root_hash.len() != 32 (guest try_into().unwrap()); host shim only */
#define MPTV_ST_DEP_FAILED 7         /* nested: the account proof this root comes from was rejected, or
                                        its value is not an Account RLP (storage-circuit main.rs:10-15) */

/* ---- error codes */
#define MPTV_OK 0
#define MPTV_ERR_ARG (-1)    /* null / inconsistent argument                                         */
#define MPTV_ERR_CUDA (-2)   /* CUDA runtime error; mptv_last_error(ctx) has the text                */
#define MPTV_ERR_ALIGN (-3)  /* a node does not start on a 16-byte boundary of the arena             */
#define MPTV_ERR_NOMEM (-4)  /* device / pinned allocation failed, or an output capacity is too small     */
#define MPTV_ERR_DEP (-5)    /* root_from_proof must point at an EARLIER, independent proof          */
#define MPTV_ERR_NODEV (-6)  /* no CUDA device / CUDA extension unusable: there is no CPU fallback   */

typedef struct mptv_ctx mptv_ctx;

/*
 * A batch in flat CSR form (what MerkleProofInput { proof, root_hash, key } x n_proofs flattens to):
 *   node i      = node_bytes[node_off[i] .. node_off[i] + node_len[i])
 *                 node_off[i] % 16 == 0, and node_bytes is readable up to the next multiple of 16
 *                 after each node (padding bytes are never hashed); node_len[i] <= 0xffff0000
 *   proof p     = nodes [proof_first[p], proof_first[p+1])          (any order, duplicates allowed)
 *   root of p   = roots[32p .. 32p+32)
 *   key of p    = key_bytes[key_off[p] .. key_off[p+1])
 *   root_from_proof (optional, may be NULL): -1, or the index d < p of an independent proof whose
 *                 returned value is an Account RLP; then p's root is that account's storage_root
 *                 and roots[32p..] is ignored (StorageProofInput, types.rs:11-19).
 */
typedef struct mptv_batch {
  const uint8_t* node_bytes;
  uint64_t node_bytes_len;
  const uint64_t* node_off;  /* [n_nodes]     */
  const uint32_t* node_len;  /* [n_nodes]     */
  uint64_t n_nodes;
  const uint32_t* proof_first; /* [n_proofs+1] */
  uint64_t n_proofs;
  const uint8_t* roots;      /* [32*n_proofs] */
  const uint8_t* key_bytes;
  const uint32_t* key_off;   /* [n_proofs+1]  */
  const int32_t* root_from_proof; /* [n_proofs] or NULL */
} mptv_batch;

/* Results.  The returned value is always a contiguous slice of one supplied node, so it is
 * reported as (offset into node_bytes, length); status != 0 gives (0, 0). */
typedef struct mptv_result {
  uint8_t* status;      /* [n_proofs] MPTV_ST_* */
  uint64_t* value_off;  /* [n_proofs] */
  uint32_t* value_len;  /* [n_proofs] */
} mptv_result;

/* device time of the last mptv_verify_batch_device / mptv_keccak256_batch_device call on a device,
 * measured with CUDA events on the stream the kernels were launched on */
typedef struct mptv_timings {
  float bin_ms;     /* K0 rate-block binning            */
  float keccak_ms;  /* K1 Keccak-256 of every node      */
  float parse_ms;   /* K2a per-node decode              */
  float walk_ms;    /* K2b walk (both waves)            */
  float total_ms;
  uint64_t n_nodes;
  uint64_t n_perm;  /* reserved, always 0: the verify entries do not count permutations (sum ceil((len+1)/136) over
                       the caller's node_len gives it; a dedup_nodes run reports its executed count below)     */
  uint32_t keccak_launches, other_launches;
  uint64_t n_unique_nodes; /* "dedup_nodes" runs: distinct nodes actually hashed (0 otherwise) ... */
  uint64_t n_unique_perm;  /* ... and their Keccak-f count                                        */
} mptv_timings;

/* device_ids == NULL && n_devices == 0: use every visible CUDA device. */
int mptv_create(const int* device_ids, int n_devices, mptv_ctx** out);
void mptv_destroy(mptv_ctx* ctx);
int mptv_device_count(const mptv_ctx* ctx);
const char* mptv_last_error(const mptv_ctx* ctx);
const char* mptv_strerror(int err);
const char* mptv_status_name(int status);

/* Host-buffer entry (what the Rust shim's verify_merkle_proofs calls): every pointer of `in` and
 * `out` is HOST memory (pinned memory from mptv_alloc_pinned makes the copies faster).  The batch
 * is cut into contiguous proof slices balanced by Keccak-f count, one slice per device of the
 * context, with no inter-device traffic; slices are pipelined in chunks (H2D / kernels / D2H on
 * per-device streams).  Blocks until `out` is filled. */
int mptv_verify_batch(mptv_ctx* ctx, const mptv_batch* in, mptv_result* out);

/* The storage guest's flow in one call (storage-circuit/src/main.rs:17-27): like mptv_verify_batch, but
 * every proof p with hash_key[p] != 0 is looked up under keccak256(key p) instead of key p -- the guest's
 * `digest_keccak(&key)` for storage slots.  The keys are hashed on the device (one K1 launch); `in` is not
 * modified.  hash_key == NULL behaves exactly like mptv_verify_batch. */
int mptv_verify_batch_hashed_keys(mptv_ctx* ctx, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out);

/* borsh blobs in, verdicts out, in ONE pipelined call (the prover's input format, prover/src/bin/main.rs:41,67):
 * blob i = blobs[blob_off[i] .. blob_off[i+1]) holds borsh(MerkleProofInput) (crypto-ops/src/types.rs:4-9).  Each
 * pipeline chunk is flattened by `n_threads` host threads (<= 0: all cores but two per device of the context) straight into page-locked staging,
 * crosses PCIe as one copy and is verified while the next chunk is being flattened, so the call runs at the
 * speed of the flattener instead of flatten + copy in series.  out->value_off[i] is the offset of the returned
 * value INSIDE `blobs` (the value is a slice of a node, the node a slice of its blob).  A blob whose root_hash
 * is not 32 bytes gets MPTV_ST_BAD_ROOT_LEN; a malformed blob fails the whole call with MPTV_ERR_ARG.  A handful of blobs
 * (<= 32 proofs, <= 40 KiB: one verify_merkle_proof call's input) is flattened on the calling thread and verified by
 * ONE kernel launch ("latency_path"), ~52 us; the same holds for mptv_verify_storage_borsh. */
int mptv_verify_borsh(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads,
                      mptv_result* out);

/* alloy_rlp::decode_exact::<Account>(value) of the storage guest (storage-circuit/src/main.rs:15) on the host: 1 and the
 * 32-byte storage_root (storage_root32 may be NULL) when `value` is exactly rlp([nonce u64, balance U256, storage_root
 * B256, code_hash B256]) in canonical form, else 0 -- the guest's unwrap() panics.  The same rule the device applies
 * before a storage proof may take its root from an account leaf; a shim needs it for inputs WITHOUT storage proofs,
 * where the guest still decodes the leaf. */
int mptv_account_storage_root(const uint8_t* value, uint32_t len, uint8_t* storage_root32);

/* The risc0 storage guest over a batch of its own inputs (circuits/risc0-storage-proof/.../storage-circuit/src/main.rs:6-31):
 * blob i holds borsh(StorageProofInput) (crypto-ops/src/types.rs:11-19).  For every input the account proof is verified
 * under address_keccak against root_hash, then storage_proofs zipped with storage_keys (the shorter list decides,
 * main.rs:18-21), each under keccak256(storage key) -- hashed on the device -- against the storage_root of the verified
 * account leaf.  Same pipeline as mptv_verify_borsh (host flatten, byte-identical nodes of a chunk cross PCIe once).
 *   proof_first[n_inputs + 1]  out: input i owns results [proof_first[i], proof_first[i+1]): its account proof first,
 *                              then its storage proofs in order.  Written by the first pass, before anything is verified.
 *   results_cap                entries the arrays of `out` hold.  Fewer than proof_first[n_inputs] (e.g. 0, to ask):
 *                              MPTV_ERR_NOMEM, with proof_first filled in and nothing verified.  A caller that built the
 *                              blobs knows the count (the sum of 1 + min(storage_proofs.len(), storage_keys.len())) and
 *                              needs one call.
 *   input_status[n_inputs]     out, may be NULL: MPTV_ST_OK (the guest commits the storage values), else the status of the
 *                              first proof that fails in the guest's order; an account leaf that is not
 *                              rlp([nonce, balance, storage_root, code_hash]) (decode_exact(..).unwrap(), main.rs:15) gives
 *                              MPTV_ST_DEP_FAILED, also for an input without storage proofs.
 * out->value_off[p] is an offset into `blobs`.  root_hash.len() != 32 gives the account proof MPTV_ST_BAD_ROOT_LEN and the
 * storage proofs MPTV_ST_DEP_FAILED; a malformed blob fails the whole call with MPTV_ERR_ARG. */
int mptv_verify_storage_borsh(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n_inputs, int n_threads,
                              uint64_t* proof_first, uint8_t* input_status, uint64_t results_cap, mptv_result* out);

/* What the host-fed entries (mptv_verify_batch, mptv_verify_borsh) moved since the context was created or the
 * counters were last reset: summed over the context's devices. */
typedef struct mptv_host_stats {
  uint64_t chunks;
  uint64_t nodes, nodes_aliased;  /* supplied nodes / those aliased to an identical node instead of copied again */
  uint64_t node_bytes_supplied;   /* padded bytes of all supplied nodes                                      */
  uint64_t node_bytes_placed;     /* ... of the nodes actually staged and copied                             */
  uint64_t h2d_bytes, d2h_bytes;  /* bytes that crossed PCIe, counted from the copies issued                 */
  /* mptv_verify_borsh, microseconds of the calling thread (summed over devices): flattening chunks, waiting for a
   * free slot / the last chunk, mapping results back to blob offsets, and the whole call */
  uint64_t flatten_us, wait_us, map_us, call_us;
  uint64_t launches;              /* kernels this library queued for those calls                             */
  uint64_t pull_chunks;           /* mptv_verify_borsh chunks whose node bytes the device fetched itself from
                                     page-locked blobs (h2d_bytes counts those bytes too)                    */
  uint64_t device_chunks;         /* mptv_verify_borsh chunks flattened on the device ("borsh_mode" 1)         */
  uint64_t index_us;              /* mptv_verify_storage_borsh: the pass over the inputs' length prefixes that runs
                                     before the stream (not part of call_us)                                 */
} mptv_host_stats;
int mptv_host_stats_get(mptv_ctx* ctx, mptv_host_stats* out, int reset);

/* Device-resident entry: every pointer is DEVICE memory on the context's device `dev_index`.
 * Asynchronous on `stream` (a cudaStream_t, NULL = the context's own stream for that device). */
int mptv_verify_batch_device(mptv_ctx* ctx, int dev_index, const mptv_batch* in, mptv_result* out,
                             void* stream);

/* digest_keccak over a whole CSR arena: digests32[32i..] = keccak256(node i).  Host buffers; the nodes are cut into
 * contiguous index ranges with equal shares of the bytes, one per device of the context (no inter-device traffic). */
int mptv_keccak256_batch(mptv_ctx* ctx, const uint8_t* node_bytes, uint64_t node_bytes_len,
                         const uint64_t* node_off, const uint32_t* node_len, uint64_t n_nodes,
                         uint8_t* digests32);
/* same, device pointers, asynchronous on `stream` */
int mptv_keccak256_batch_device(mptv_ctx* ctx, int dev_index, const uint8_t* node_bytes,
                                const uint64_t* node_off, const uint32_t* node_len, uint64_t n_nodes,
                                uint8_t* digests32, void* stream);

/* synchronises the device's stream and reports the device times of its last *_device call */
int mptv_last_timings(mptv_ctx* ctx, int dev_index, mptv_timings* out);

/* Integer issue-rate probe, the denominator of the Keccak roofline: runs a LOP3 (mode 0), SHF (mode 1)
 * or Keccak-mix 122:58 (mode 2) kernel on every SM and reports 32-bit lane-operations per second. */
int mptv_int_issue_peak(mptv_ctx* ctx, int dev_index, int mode, double* lane_ops_per_s);

/* options (name, value):
 *   "lanes_per_proof"  K2b lanes per proof: 0 = choose from nodes/proof, else 8, 16 or 32
 *   "chunk_bytes"      node bytes per pipeline chunk of the host-buffer entry (default 96 MiB)
 *   "borsh_chunk_bytes" borsh bytes per pipeline chunk of mptv_verify_borsh (default 32 MiB; the device pipeline of
 *                      borsh_mode 1 / 2 takes chunks of twice this)
 *   "borsh_mode"       mptv_verify_borsh: -1 (default) = 0 when the context has one device, 1 when it drives several;
 *                      0 = the HOST flattens: a pool of threads reads every blob once and
 *                      stages the nodes, byte-identical nodes of a chunk only once ("host_dedup") -- fewest PCIe bytes,
 *                      the right mode when one GPU has the host to itself; 1 = the DEVICE flattens: the blobs, which
 *                      must be in page-locked memory, cross PCIe as they are and kernels lay the nodes out -- the cores
 *                      touch nothing, the right mode when several GPUs share one host's memory system; 2 = BOTH at once
 *                      on each device: the host pipeline takes chunks from the front of the device's blob range, the
 *                      device pipeline from its back, until they meet -- one is bound by the cores, the other by PCIe
 *                      (1 and 2 fall back to 0 for pageable blobs).  Measured per million config-2 proofs (DESIGN.md
 *                      sections 6, 7): one GPU 50 / 62 / 50-59 ms in mode 0 / 1 / 2; eight GPUs behind one host 25 M
 *                      proofs/s in mode 0, 54 M in mode 1.  Results are identical in all three
 *   "hybrid_device_pct" borsh_mode 2: the share (1 ... 100 %, default 24) of the bytes the device pipeline may take; it
 *                      is bound by PCIe and unthrottled would starve the host pipeline's copies
 *   "pull_pinned"      mptv_verify_borsh, blobs in page-locked memory (mptv_alloc_pinned, cudaHostAlloc, cudaHostRegister):
 *                      the staging copy carries only the index arrays and a gather list, and a kernel fetches the node
 *                      bytes straight from the blobs over PCIe -- the cores do not copy them and the DMA engine does not
 *                      read them a second time.  Default 0: on the measured hosts the kernel's small PCIe reads reach
 *                      30 GB/s against the copy engine's 55 and slow the flattening threads down (71 vs 50 ms per
 *                      million config-2 proofs, profiles/r02_borsh_pull_quick.txt)
 *   "wc_staging"       mptv_verify_borsh / mptv_verify_storage_borsh stage the node bytes of a chunk in write-combining
 *                      page-locked memory (default 0: no measurable gain on the measured hosts, profiles/r02_wc_staging_probe.txt)
 *   "host_dedup"       mptv_verify_borsh aliases byte-identical nodes of a chunk instead of staging and copying them
 *                      again (default 1).  Transfer de-duplication only: every supplied node is still hashed on the
 *                      device, results are identical
 *   "latency_path"     a call whose (chunk of the) batch fits one CTA -- at most 128 nodes, 32 proofs, 64 KiB packed;
 *                      every single verify_merkle_proof call does -- is verified by ONE kernel launch that reads the
 *                      inputs from, and writes the results to, mapped page-locked memory (default 1)
 *   "binning"          K0 rate-block binning on / off              (default 1)
 *   "fused_classify"   K1 also classifies plain branches / leaves  (default 1)
 *   "fast_walk"        K2f thread-per-proof chain check + K2b on the deferred rest (default 1)
 *   "fused_leaf_hash"  rebuild: K1L hashes leaves straight from the value arena (default 1)
 *   "long_leaf_bin", "long_leaf_ctas"  rebuild: leaves of at least this many rate blocks are hashed in a launch of
 *                      their own with this many 4-warp CTAs per SM (defaults 33 and 1; tuning knobs)
 *   "l2_fetch_granularity"  32, 64 or 128: device-wide cudaLimitMaxL2FetchGranularity hint for the context's GPUs
 *                      (left at the driver default unless set; a tuning knob)
 *   "dedup_nodes"      hash each DISTINCT node of a batch once and share the digest (default 0).
 *                      Results are identical; the executed Keccak-f count drops.  The reference hashes
 *                      every supplied node, so numbers measured with this option are a SECOND,
 *                      separately labelled figure, never the headline. */
int mptv_set_option(mptv_ctx* ctx, const char* name, int64_t value);

/* ---------------------------------------------------------------------------------------------
 * Trie rebuild: what trie-utils does per block with EthTrie::new + insert(rlp(i), bytes) x n +
 * root_hash()  (/root/reference/trie-utils/src/proofs/transaction.rs:41-66, proofs/receipt.rs:49-84,
 * src/receipt.rs:8-38), for a whole batch of independent tries at once.
 *   item i      = key   key_bytes[key_off[i] .. key_off[i+1])            (<= 32 bytes)
 *                 value value_bytes[value_off[i] .. value_off[i] + value_len[i])
 *                 value_off[i] % 16 == 0, value_bytes readable up to the next multiple of 16 after
 *                 each value; value_len[i] == 0 deletes the key (eth_trie: insert(k, b"") == remove)
 *   trie t      = items [trie_first[t], trie_first[t+1]) applied in order (last write wins);
 *                 at most 8192 items per trie (per-block tries; larger ones are refused: MPTV_ERR_ARG)
 *   roots32     = [32 * n_tries]; an empty trie gives keccak256(0x80).
 */
typedef struct mptv_kv_batch {
  const uint8_t* key_bytes;
  const uint32_t* key_off;    /* [n_items+1] */
  const uint8_t* value_bytes;
  uint64_t value_bytes_len;
  const uint64_t* value_off;  /* [n_items]   */
  const uint32_t* value_len;  /* [n_items]   */
  uint64_t n_items;
  const uint32_t* trie_first; /* [n_tries+1] */
  uint64_t n_tries;
} mptv_kv_batch;

/* device time of the last mptv_trie_roots_device call on a device (CUDA events on its stream) */
typedef struct mptv_rebuild_timings {
  float structure_ms; /* sort + skeleton + level lists (incl. the summary read-back)     */
  float encode_ms;    /* sum over levels of the RLP encode launches                      */
  float keccak_ms;    /* sum over levels of K0 + K1                                      */
  float total_ms;
  uint64_t n_nodes;   /* trie nodes built                                                */
  uint64_t n_hashed;  /* nodes >= 32 bytes (+ roots): the ones write_node hashes         */
  uint64_t n_perm;    /* algorithmic Keccak-f count = sum over hashed nodes ceil((len+1)/136) */
  uint64_t arena_bytes;
  uint32_t levels, keccak_launches, other_launches, pad;
} mptv_rebuild_timings;

/* host buffers; tries are sharded over the context's devices as contiguous slices balanced by
 * value bytes, no inter-device traffic */
int mptv_trie_roots(mptv_ctx* ctx, const mptv_kv_batch* in, uint8_t* roots32);
/* device pointers on device `dev_index`; runs on `stream` (NULL = the context's stream) and
 * returns after the launches of the last level are queued (it synchronises the stream twice in
 * between to read back the level sizes) */
int mptv_trie_roots_device(mptv_ctx* ctx, int dev_index, const mptv_kv_batch* in, uint8_t* roots32, void* stream);
int mptv_last_rebuild_timings(mptv_ctx* ctx, int dev_index, mptv_rebuild_timings* out);

/* Rebuild + Trie::get_proof: what get_ethereum_transaction_proof_inputs / get_ethereum_receipt_proof_inputs
 * do after fetching the block (/root/reference/trie-utils/src/proofs/transaction.rs:41-73,
 * proofs/receipt.rs:49-92): build the tries, take root_hash(), and for every target (trie, key) emit
 * the encoded nodes on the key's path, root first (the root always, other nodes only when referenced
 * by hash; for an absent key the path that proves its absence).  The output is laid out as an
 * mptv_batch arena (every node on a 16-byte boundary), so {proof q, roots32[trie[q]], key q} can be
 * handed straight to mptv_verify_batch.  Host buffers.  When the targets are grouped by trie in ascending order (the
 * usual shape) the tries and their targets are cut into contiguous slices over the devices of the context, like
 * mptv_trie_roots; otherwise device 0 does the call. */
typedef struct mptv_proof_targets {
  const uint32_t* trie;      /* [n_targets] trie the key is looked up in */
  const uint8_t* key_bytes;  /* target q key = key_bytes[key_off[q] .. key_off[q+1]) */
  const uint32_t* key_off;   /* [n_targets+1] */
  uint64_t n_targets;
} mptv_proof_targets;
typedef struct mptv_proofs_out {
  uint8_t* node_bytes;       /* caller-allocated, node_bytes_cap bytes */
  uint64_t node_bytes_cap;
  uint64_t* node_off;        /* [nodes_cap] */
  uint32_t* node_len;        /* [nodes_cap] */
  uint64_t nodes_cap;
  uint32_t* proof_first;     /* [n_targets+1] */
  uint64_t n_nodes;          /* out: nodes written -- or required, when MPTV_ERR_NOMEM is returned */
  uint64_t node_bytes_len;   /* out: bytes written -- or required */
} mptv_proofs_out;
int mptv_trie_proofs(mptv_ctx* ctx, const mptv_kv_batch* in, const mptv_proof_targets* targets, uint8_t* roots32,
                     mptv_proofs_out* out);

/* ---------------------------------------------------------------------------------------------
 * Host-side formats in front of the hot path (no GPU needed for these).
 *
 * mptv_flatten_borsh: n borsh-serialised MerkleProofInput blobs (types.rs:4-9; blob i =
 * blobs[blob_off[i] .. blob_off[i+1])) -> one CSR batch, multi-threaded (n_threads <= 0: all
 * cores), in page-locked memory when `pinned`.  Malformed borsh (what borsh::from_slice rejects)
 * gives MPTV_ERR_ARG.  *out must be NULL, or a handle returned by an earlier call: its buffers are then
 * recycled (grown when needed), so a steady-state pipeline pays for page faults / page-locking once.
 * A root_hash that is not 32 bytes is flagged in bad_root (the guests'
 * try_into().unwrap() panic, MPTV_ST_BAD_ROOT_LEN) and its proof is verified against a zero root. */
typedef struct mptv_host_batch mptv_host_batch;
int mptv_flatten_borsh(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                       mptv_host_batch** out);
/* Same input and output, in ONE pass over the blobs, with flags:
 *   MPTV_FLATTEN_ALIAS_DUPLICATES  a node that is byte-identical (exact compare) to an earlier node of the input is
 *       not stored again: its node_off points at the first copy.  Proofs against one trie share their upper nodes, so
 *       the arena shrinks several-fold (config 2: 3.1 GB -> ~0.5 GB) and so does what mptv_verify_batch has to move.
 *       Verification results are identical: the device still hashes every supplied node.
 * The batch's node_bytes block also holds the index arrays (node_off values are offsets into that block); nodes are
 * NOT laid out in index order.  info (may be NULL) reports what was shared. */
#define MPTV_FLATTEN_ALIAS_DUPLICATES 1u
typedef struct mptv_flatten_info {
  uint64_t n_nodes, nodes_aliased;
  uint64_t node_bytes_supplied; /* sum of padded node lengths */
  uint64_t node_bytes_placed;   /* ... of the nodes actually stored */
} mptv_flatten_info;
int mptv_flatten_borsh_ex(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                          unsigned flags, mptv_host_batch** out, mptv_flatten_info* info);
/* The same for borsh(StorageProofInput) blobs (crypto-ops/src/types.rs:11-19): each input becomes its account proof
 * (key = address_keccak, root = root_hash) followed by min(storage_proofs.len(), storage_keys.len()) storage proofs
 * (key = the RAW storage key, root_from_proof = the account proof), i.e. the batch mptv_verify_batch_hashed_keys takes
 * together with *hash_key (one flag per proof, owned by the handle): the host half of mptv_verify_storage_borsh.
 * proof_first [n_inputs + 1] as there. */
int mptv_flatten_storage_borsh(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n_inputs, int n_threads, int pinned,
                               unsigned flags, mptv_host_batch** out, mptv_flatten_info* info, uint64_t* proof_first,
                               const uint8_t** hash_key);
/* The host stage of mptv_verify_borsh alone (no device needed): the same chunk loop and builder into ordinary memory.
 * Timed by the caller, it is the ceiling of the streamed entry on this host with these threads. */
int mptv_borsh_flatten_probe(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, uint64_t chunk_bytes,
                             int alias_duplicates, mptv_flatten_info* info);
/* What the host's memory system gives n_threads threads that each stream a private buffer of bytes_per_thread:
 * read-only GB/s, and GB/s of payload for a copy with non-temporal stores (the memory traffic is twice that).  The
 * ceilings of the host stage above. */
int mptv_host_bw_probe(int n_threads, uint64_t bytes_per_thread, double* read_gbs, double* copy_gbs);
const mptv_batch* mptv_host_batch_view(const mptv_host_batch* hb);
const uint8_t* mptv_host_batch_bad_root(const mptv_host_batch* hb); /* [n_proofs] */
void mptv_host_batch_free(mptv_host_batch* hb);

/* alloy_rlp::encode(index): the trie key of transaction / receipt `index` (transaction.rs:45); returns the length */
uint32_t mptv_rlp_index(uint64_t index, uint8_t out[9]);

/* trie-utils insert_receipt's leaf bytes (receipt.rs:8-38): [prefix] ++ rlp([status, cumulative_gas_used,
 * bloom, logs]); prefix < 0 = legacy receipt (no type byte).  Returns the encoded length; writes
 * min(length, cap) bytes (call with out == NULL to size). */
typedef struct mptv_log {
  const uint8_t* address;  /* 20 bytes */
  const uint8_t* topics;   /* 32 * n_topics bytes */
  uint32_t n_topics;
  const uint8_t* data;
  uint32_t data_len;
} mptv_log;
uint64_t mptv_encode_receipt(int prefix, int status, uint64_t cumulative_gas_used, const uint8_t* bloom256,
                             const mptv_log* logs, uint32_t n_logs, uint8_t* out, uint64_t cap);

/* ---- ABI layout pins.  Bindings mirror these structures field by field (Rust #[repr(C)] in
 * integration/rust/crypto-ops-gpu/src/lib.rs, ctypes.Structure in zk-state-proofs_b200/crypto_ops.py); a change of
 * size or field offset on the LP64 targets this library is built for must fail the build here, not corrupt a call. */
#define MPTV_ABI_PIN(name, cond) typedef char mptv_abi_pin_##name[(cond) ? 1 : -1]
MPTV_ABI_PIN(batch_size, sizeof(mptv_batch) == 88);
MPTV_ABI_PIN(batch_node_bytes_len, offsetof(mptv_batch, node_bytes_len) == 8);
MPTV_ABI_PIN(batch_node_off, offsetof(mptv_batch, node_off) == 16);
MPTV_ABI_PIN(batch_node_len, offsetof(mptv_batch, node_len) == 24);
MPTV_ABI_PIN(batch_n_nodes, offsetof(mptv_batch, n_nodes) == 32);
MPTV_ABI_PIN(batch_proof_first, offsetof(mptv_batch, proof_first) == 40);
MPTV_ABI_PIN(batch_n_proofs, offsetof(mptv_batch, n_proofs) == 48);
MPTV_ABI_PIN(batch_roots, offsetof(mptv_batch, roots) == 56);
MPTV_ABI_PIN(batch_key_bytes, offsetof(mptv_batch, key_bytes) == 64);
MPTV_ABI_PIN(batch_key_off, offsetof(mptv_batch, key_off) == 72);
MPTV_ABI_PIN(batch_root_from_proof, offsetof(mptv_batch, root_from_proof) == 80);
MPTV_ABI_PIN(result_size, sizeof(mptv_result) == 24);
MPTV_ABI_PIN(result_value_off, offsetof(mptv_result, value_off) == 8);
MPTV_ABI_PIN(result_value_len, offsetof(mptv_result, value_len) == 16);
MPTV_ABI_PIN(kv_batch_size, sizeof(mptv_kv_batch) == 72);
MPTV_ABI_PIN(kv_value_bytes_len, offsetof(mptv_kv_batch, value_bytes_len) == 24);
MPTV_ABI_PIN(kv_n_items, offsetof(mptv_kv_batch, n_items) == 48);
MPTV_ABI_PIN(kv_n_tries, offsetof(mptv_kv_batch, n_tries) == 64);
MPTV_ABI_PIN(proof_targets_size, sizeof(mptv_proof_targets) == 32);
MPTV_ABI_PIN(proofs_out_size, sizeof(mptv_proofs_out) == 64);
MPTV_ABI_PIN(proofs_out_n_nodes, offsetof(mptv_proofs_out, n_nodes) == 48);
MPTV_ABI_PIN(timings_size, sizeof(mptv_timings) == 64);
MPTV_ABI_PIN(rebuild_timings_size, sizeof(mptv_rebuild_timings) == 64);
MPTV_ABI_PIN(host_stats_size, sizeof(mptv_host_stats) == 120);
MPTV_ABI_PIN(flatten_info_size, sizeof(mptv_flatten_info) == 32);
MPTV_ABI_PIN(log_size, sizeof(mptv_log) == 40);
#undef MPTV_ABI_PIN

/* page-locked host memory for arenas that are handed to mptv_verify_batch */
void* mptv_alloc_pinned(size_t bytes);
void mptv_free_pinned(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MPTV_H */
