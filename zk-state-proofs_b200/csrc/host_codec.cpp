// host_codec.cpp -- host-side data formats either side of the hot path, in C++
// because the reference's host side is compiled code:
//   * borsh(MerkleProofInput) blobs -> the CSR arena of include/mptv.h, multi-threaded, two passes
//     (size, then copy), every node on a 16-byte boundary, optionally straight into page-locked memory.
//     Wire format: /root/reference/crypto-ops/src/types.rs:4-9 (derive BorshSerialize: Vec<Vec<u8>>,
//     Vec<u8>, Vec<u8> with u32-LE length prefixes), produced by the prover
//     (/root/reference/prover/src/bin/main.rs:41,67) and decoded by the guests
//     (/root/reference/circuits/sp1-merkle-proof/src/main.rs:5-6).
//   * the receipt leaf encoder of trie-utils: [prefix] ++ rlp([status, cumulative_gas_used, bloom, logs])
//     (/root/reference/trie-utils/src/receipt.rs:8-38, Log: src/types.rs:11-35; known answer:
//     /root/reference/trie-utils/tests/rlp.rs:12) and alloy_rlp::encode(index), the trie key
//     (/root/reference/trie-utils/src/proofs/transaction.rs:45).
// No hashing, trie walking or RLP *decoding* happens here: this is layout work in front of the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <new>
#include <thread>
#include <vector>

#include "../../include/mptv.h"
#include "host_codec.h"
#include "host_flatten.h"

using mptv::BlobShape;
using mptv::borsh_shape;
using mptv::parallel_for;

struct mptv_host_batch {
  mptv_batch view;
  uint8_t* bad_root = nullptr;  // [n] 1 where root_hash.len() != 32
  bool pinned = false;
  // the eight arrays of the batch; kept (and grown) across calls when the handle is reused, so a
  // steady-state pipeline pays for page faults / page-locking once
  void* blocks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

void* host_alloc(mptv_host_batch* hb, int slot, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (hb->blocks[slot] && hb->cap[slot] >= bytes) return hb->blocks[slot];
  if (hb->blocks[slot]) {
    if (hb->pinned) cudaFreeHost(hb->blocks[slot]); else free(hb->blocks[slot]);
    hb->blocks[slot] = nullptr; hb->cap[slot] = 0;
  }
  void* p = nullptr;
  const size_t want = bytes + bytes / 8;
  if (hb->pinned) {
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) p = nullptr;
  } else {
    if (posix_memalign(&p, 64, want) != 0) p = nullptr;
  }
  if (p) { hb->blocks[slot] = p; hb->cap[slot] = want; }
  return p;
}

// ---- RLP writers (alloy-rlp Encodable)
struct Out {
  uint8_t* p; uint64_t cap; uint64_t n = 0;
  void put(uint8_t b) { if (n < cap) p[n] = b; n++; }
  void put(const uint8_t* s, uint64_t k) {
    if (n + k <= cap) memcpy(p + n, s, k);
    else if (n < cap) memcpy(p + n, s, cap - n);
    n += k;
  }
};
inline uint32_t be_len(uint64_t v) { uint32_t k = 0; while (v) { k++; v >>= 8; } return k; }
inline uint64_t hdr_len(uint64_t payload) { return payload < 56 ? 1 : 1 + be_len(payload); }
void put_hdr(Out& o, uint64_t payload, bool list) {
  const uint8_t base = list ? 0xC0 : 0x80;
  if (payload < 56) { o.put((uint8_t)(base + payload)); return; }
  const uint32_t k = be_len(payload);
  o.put((uint8_t)(base + 55 + k));
  for (uint32_t i = 0; i < k; i++) o.put((uint8_t)(payload >> (8 * (k - 1 - i))));
}
inline uint64_t str_len(const uint8_t* s, uint64_t n) { return (n == 1 && s[0] < 0x80) ? 1 : hdr_len(n) + n; }
void put_str(Out& o, const uint8_t* s, uint64_t n) {
  if (n == 1 && s[0] < 0x80) { o.put(s[0]); return; }
  put_hdr(o, n, false);
  o.put(s, n);
}
inline uint64_t u64_len(uint64_t v) { return v < 0x80 ? 1 : 1 + be_len(v); }
void put_u64(Out& o, uint64_t v) {
  if (v == 0) { o.put(0x80); return; }
  if (v < 0x80) { o.put((uint8_t)v); return; }
  const uint32_t k = be_len(v);
  o.put((uint8_t)(0x80 + k));
  for (uint32_t i = 0; i < k; i++) o.put((uint8_t)(v >> (8 * (k - 1 - i))));
}
inline uint64_t log_payload(const mptv_log& l) {
  const uint64_t topics = 33ull * l.n_topics;
  return 21 + hdr_len(topics) + topics + str_len(l.data, l.data_len);
}

}  // namespace

extern "C" {

static int flatten_borsh_run(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                             mptv_host_batch** out) {
  if (!out || (n && (!blobs || !blob_off))) return MPTV_ERR_ARG;
  mptv_host_batch* reuse = *out;  // NULL, or a handle from an earlier call whose buffers are recycled
  if (reuse && reuse->pinned != (pinned != 0)) return MPTV_ERR_ARG;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<BlobShape> sh(n);
  std::atomic<int> bad(0);
  parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      if (blob_off[i + 1] < blob_off[i]) { bad = 1; continue; }
      sh[i] = borsh_shape(blobs + blob_off[i], blobs + blob_off[i + 1]);
      if (!sh[i].ok) bad = 1;
    }
  });
  if (bad) return MPTV_ERR_ARG;
  // exclusive scans (serial: three adds per input)
  std::vector<uint64_t> node_first(n + 1, 0), byte_first(n + 1, 0), key_first(n + 1, 0);
  for (uint64_t i = 0; i < n; i++) {
    node_first[i + 1] = node_first[i] + sh[i].n_nodes;
    byte_first[i + 1] = byte_first[i] + sh[i].padded_bytes;
    key_first[i + 1] = key_first[i] + sh[i].key_len;
  }
  if (node_first[n] > 0xfffffff0ull || key_first[n] > 0xfffffff0ull) return MPTV_ERR_ARG;
  mptv_host_batch* hb = reuse ? reuse : new mptv_host_batch();
  hb->pinned = pinned != 0;
  const uint64_t nn = node_first[n], nb = byte_first[n] + 16;
  uint8_t* node_bytes = (uint8_t*)host_alloc(hb, 0, nb);
  uint64_t* node_off = (uint64_t*)host_alloc(hb, 1, 8 * nn);
  uint32_t* node_len = (uint32_t*)host_alloc(hb, 2, 4 * nn);
  uint32_t* proof_first = (uint32_t*)host_alloc(hb, 3, 4 * (n + 1));
  uint8_t* roots = (uint8_t*)host_alloc(hb, 4, 32 * n);
  uint8_t* key_bytes = (uint8_t*)host_alloc(hb, 5, key_first[n] + 16);
  uint32_t* key_off = (uint32_t*)host_alloc(hb, 6, 4 * (n + 1));
  hb->bad_root = (uint8_t*)host_alloc(hb, 7, n);
  if (!node_bytes || !node_off || !node_len || !proof_first || !roots || !key_bytes || !key_off || !hb->bad_root) {
    mptv_host_batch_free(hb);
    *out = nullptr;
    return MPTV_ERR_NOMEM;
  }
  memset(node_bytes + byte_first[n], 0, 16);
  memset(key_bytes + key_first[n], 0, 16);
  proof_first[n] = (uint32_t)nn;
  key_off[n] = (uint32_t)key_first[n];
  // pass 2: copy
  parallel_for(n, n_threads, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      proof_first[i] = (uint32_t)node_first[i];
      key_off[i] = (uint32_t)key_first[i];
      hb->bad_root[i] = sh[i].bad_root;
      mptv::borsh_copy(blobs + blob_off[i], sh[i], node_bytes, byte_first[i], node_off, node_len, node_first[i],
                       roots + 32 * i, key_bytes + key_first[i], nullptr, blobs);
    }
    _mm_sfence();  // the nodes were written with non-temporal stores
  });
  hb->view.node_bytes = node_bytes; hb->view.node_bytes_len = nb;
  hb->view.node_off = node_off; hb->view.node_len = node_len; hb->view.n_nodes = nn;
  hb->view.proof_first = proof_first; hb->view.n_proofs = n; hb->view.roots = roots;
  hb->view.key_bytes = key_bytes; hb->view.key_off = key_off; hb->view.root_from_proof = nullptr;
  *out = hb;
  return MPTV_OK;
}

int mptv_flatten_borsh(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                       mptv_host_batch** out) {
  try {  // the C ABI never throws: the per-input tables below are sized by the caller's n
    return flatten_borsh_run(blobs, blob_off, n, n_threads, pinned, out);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// Single pass + optional aliasing of byte-identical nodes (host_flatten.h).  Block 0 of the handle holds the index
// arrays and the workers' byte regions; the keys are re-packed into the public prefix-array form (32 bytes a proof).
// storage: the blobs are borsh(StorageProofInput); proof_first [n + 1] and *hash_key (one flag per proof, inside the
// handle's block) are written as well, and the batch carries root_from_proof
static int flatten_borsh_ex_run(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                                unsigned flags, mptv_host_batch** out, mptv_flatten_info* info, bool storage = false,
                                uint64_t* proof_first = nullptr, const uint8_t** hash_key = nullptr) {
  if (!out || (n && (!blobs || !blob_off))) return MPTV_ERR_ARG;
  if (storage && (!proof_first || !hash_key)) return MPTV_ERR_ARG;
  if (flags & ~(unsigned)MPTV_FLATTEN_ALIAS_DUPLICATES) return MPTV_ERR_ARG;
  mptv_host_batch* reuse = *out;
  if (reuse && reuse->pinned != (pinned != 0)) return MPTV_ERR_ARG;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (n < 256) n_threads = 1;
  mptv::WorkerPool pool(n_threads);
  mptv::DedupTable table;
  const bool alias = (flags & MPTV_FLATTEN_ALIAS_DUPLICATES) != 0;
  if (alias) {
    // one chunk = the whole input: a node worth sharing is >= 128 bytes, so bytes / 128 bounds the entries
    const uint64_t bytes = n ? blob_off[n] - blob_off[0] : 0;
    if (!table.reserve((size_t)std::min<uint64_t>(bytes / 256 + 1024, 1ull << 24))) return MPTV_ERR_NOMEM;
    table.new_epoch();
  }
  mptv_host_batch* hb = reuse ? reuse : new mptv_host_batch();
  hb->pinned = pinned != 0;
  auto fail = [&](int rc) {
    if (!reuse) mptv_host_batch_free(hb);
    return rc;
  };
  mptv::ChunkLayout L;
  std::vector<uint8_t> bad;
  auto get_block = [&](size_t total, size_t) { return (uint8_t*)host_alloc(hb, 0, total); };
  int rc;
  uint64_t n_in = n;  // inputs; n becomes the number of proofs
  if (storage) {
    mptv::StorageIndex idx;
    rc = mptv::skim_storage_inputs(pool, blobs, blob_off, n, idx);
    if (rc != MPTV_OK) return fail(rc);
    memcpy(proof_first, idx.proof_first.data(), 8 * (n + 1));
    mptv::StorageChunkJob job = {blobs, blob_off, 0, n, alias ? &table : nullptr, &idx, /*key_off_16=*/true};
    rc = mptv::flatten_storage_chunk(pool, job, get_block, L, nullptr, &bad);
    n = idx.proof_first[n_in];
  } else {
    mptv::BorshChunkJob job = {blobs, blob_off, 0, n, alias ? &table : nullptr, false, /*key_off_16=*/true};
    rc = mptv::flatten_borsh_chunk(pool, job, get_block, L, nullptr, &bad);
  }
  if (rc != MPTV_OK) return fail(rc);
  (void)n_in;
  uint8_t* block = (uint8_t*)hb->blocks[0];
  const uint32_t* koff = reinterpret_cast<const uint32_t*>(block + L.o_koff);
  const uint32_t* klen = reinterpret_cast<const uint32_t*>(block + L.o_klen);
  uint64_t kbytes = 0;
  for (uint64_t i = 0; i < n; i++) kbytes += klen[i];
  if (kbytes > 0xfffffff0ull) return fail(MPTV_ERR_ARG);
  uint8_t* key_bytes = (uint8_t*)host_alloc(hb, 5, kbytes + 16);
  uint32_t* key_off = (uint32_t*)host_alloc(hb, 6, 4 * (n + 1));
  hb->bad_root = (uint8_t*)host_alloc(hb, 7, n);
  if (!key_bytes || !key_off || !hb->bad_root) return fail(MPTV_ERR_NOMEM);
  uint64_t o = 0;
  for (uint64_t i = 0; i < n; i++) {
    key_off[i] = (uint32_t)o;
    if (klen[i]) memcpy(key_bytes + o, block + ((uint64_t)koff[i] << 4), klen[i]);
    o += klen[i];
    hb->bad_root[i] = bad[i];
  }
  key_off[n] = (uint32_t)o;
  memset(key_bytes + o, 0, 16);
  hb->view.node_bytes = block; hb->view.node_bytes_len = L.total;
  hb->view.node_off = reinterpret_cast<const uint64_t*>(block + L.o_off);
  hb->view.node_len = reinterpret_cast<const uint32_t*>(block + L.o_len);
  hb->view.n_nodes = L.nn;
  hb->view.proof_first = reinterpret_cast<const uint32_t*>(block + L.o_pf);
  hb->view.n_proofs = n; hb->view.roots = block + L.o_roots;
  hb->view.key_bytes = key_bytes; hb->view.key_off = key_off; hb->view.root_from_proof = nullptr;
  if (storage) {
    hb->view.root_from_proof = reinterpret_cast<const int32_t*>(block + L.o_rfp);
    *hash_key = block + L.o_hk;
  }
  if (info) {
    info->n_nodes = L.nn; info->nodes_aliased = L.nodes_aliased;
    info->node_bytes_supplied = L.node_bytes_supplied; info->node_bytes_placed = L.node_bytes_placed;
  }
  *out = hb;
  return MPTV_OK;
}

int mptv_flatten_borsh_ex(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, int pinned,
                          unsigned flags, mptv_host_batch** out, mptv_flatten_info* info) {
  try {
    return flatten_borsh_ex_run(blobs, blob_off, n, n_threads, pinned, flags, out, info);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// strict canonical RLP header (alloy-rlp Header::decode), as rlp_hdr of verify_device.cuh
static bool host_rlp_hdr(const uint8_t* p, uint32_t n, bool& is_list, uint32_t& hdr_len, uint32_t& payload_len) {
  if (n == 0) return false;
  const uint32_t b = p[0];
  if (b < 0x80) { is_list = false; hdr_len = 0; payload_len = 1; return true; }
  if (b < 0xB8) {
    is_list = false; hdr_len = 1; payload_len = b - 0x80;
    if (payload_len == 1 && (n < 2 || p[1] < 0x80)) return false;
  } else if (b < 0xC0 || b >= 0xF8) {
    is_list = b >= 0xF8;
    const uint32_t ll = is_list ? b - 0xF7 : b - 0xB7;
    if (n < 1 + ll || p[1] == 0 || ll > 4) return false;
    uint32_t v = 0;
    for (uint32_t i = 0; i < ll; i++) v = (v << 8) | p[1 + i];
    if (v < 56) return false;
    hdr_len = 1 + ll; payload_len = v;
  } else {
    is_list = true; hdr_len = 1; payload_len = b - 0xC0;
  }
  return (uint64_t)hdr_len + payload_len <= (uint64_t)n;
}

// alloy_rlp::decode_exact::<Account> (storage-circuit/src/main.rs:15), the rule set of account_storage_root_off on the
// device: rlp([nonce u64, balance U256, storage_root B256, code_hash B256]) and nothing else
int mptv_account_storage_root(const uint8_t* v, uint32_t n, uint8_t* storage_root32) {
  if (!v) return 0;
  bool lst, tl;
  uint32_t hl, pl, thl, tpl, at = 0;
  if (!host_rlp_hdr(v, n, lst, hl, pl) || !lst || hl + pl != n) return 0;
  uint32_t q = hl;
  for (int i = 0; i < 4; i++) {
    if (!host_rlp_hdr(v + q, n - q, tl, thl, tpl) || tl) return 0;
    if (i == 0 && tpl > 8) return 0;
    if (i == 1 && tpl > 32) return 0;
    if (i < 2 && tpl > 0 && v[q + thl] == 0) return 0;
    if (i >= 2 && tpl != 32) return 0;
    if (i == 2) at = q + thl;
    q += thl + tpl;
  }
  if (q != n) return 0;
  if (storage_root32) memcpy(storage_root32, v + at, 32);
  return 1;
}

int mptv_flatten_storage_borsh(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n_inputs, int n_threads, int pinned,
                               unsigned flags, mptv_host_batch** out, mptv_flatten_info* info, uint64_t* proof_first,
                               const uint8_t** hash_key) {
  try {
    return flatten_borsh_ex_run(blobs, blob_off, n_inputs, n_threads, pinned, flags, out, info, true, proof_first, hash_key);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// The host stage of mptv_verify_borsh run ALONE (no device, ordinary memory): the same chunk loop, the same
// builder, three recycled blocks.  Its rate is the ceiling of the streamed entry on this host.
int mptv_borsh_flatten_probe(const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, uint64_t chunk_bytes,
                             int alias_duplicates, mptv_flatten_info* info) {
  if (n && (!blobs || !blob_off)) return MPTV_ERR_ARG;
  try {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (chunk_bytes < 4096) chunk_bytes = 32ull << 20;
    mptv::WorkerPool pool(n_threads);
    mptv::DedupTable table;
    if (alias_duplicates && !table.reserve((size_t)std::min<uint64_t>(chunk_bytes / 128 + 1024, 1ull << 22))) return MPTV_ERR_NOMEM;
    // recycled across calls (like the page-locked staging of the real entry): page faults are paid once
    struct Blk { uint8_t* p = nullptr; size_t cap = 0; };
    static thread_local Blk block[3];
    mptv_flatten_info acc = {0, 0, 0, 0};
    size_t ci = 0;
    for (uint64_t cs = 0; cs < n; ci++) {
      uint64_t ce = (uint64_t)(std::upper_bound(blob_off + cs + 1, blob_off + n + 1, blob_off[cs] + chunk_bytes) - blob_off);
      if (ce > cs + 1) ce--;
      if (ce > n) ce = n;
      table.new_epoch();
      mptv::BorshChunkJob job = {blobs, blob_off, cs, ce, alias_duplicates ? &table : nullptr, false};
      mptv::ChunkLayout L;
      Blk& blk = block[ci % 3];
      const int rc = mptv::flatten_borsh_chunk(pool, job, [&](size_t total, size_t) {
        if (blk.cap < total) {
          free(blk.p);
          blk.cap = total + total / 8;
          if (posix_memalign((void**)&blk.p, 64, blk.cap) != 0) { blk.p = nullptr; blk.cap = 0; }
        }
        return blk.p;
      }, L, nullptr, nullptr);
      if (rc != MPTV_OK) return rc;
      acc.n_nodes += L.nn; acc.nodes_aliased += L.nodes_aliased;
      acc.node_bytes_supplied += L.node_bytes_supplied; acc.node_bytes_placed += L.node_bytes_placed;
      cs = ce;
    }
    if (info) *info = acc;
    return MPTV_OK;
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// What the host's memory system gives `n_threads` threads, each streaming its own buffer: reads only, and reads +
// non-temporal writes (a copy).  The ceilings the streamed entry's host stage is measured against.
int mptv_host_bw_probe(int n_threads, uint64_t bytes_per_thread, double* read_gbs, double* copy_gbs) {
  if (!read_gbs || !copy_gbs) return MPTV_ERR_ARG;
  try {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const size_t n = (size_t)((bytes_per_thread < (1u << 20) ? (1u << 20) : bytes_per_thread) & ~(uint64_t)63);
    std::vector<uint8_t*> src(n_threads, nullptr), dst(n_threads, nullptr);
    std::vector<uint64_t> sink(n_threads, 0);
    mptv::WorkerPool pool(n_threads);
    std::atomic<int> bad(0);
    pool.run([&](int t) {  // first touch by the thread that streams the buffer
      if (posix_memalign((void**)&src[t], 64, n) != 0 || posix_memalign((void**)&dst[t], 64, n) != 0) { bad = 1; return; }
      memset(src[t], t + 1, n);
      memset(dst[t], 0, n);
    });
    double tr = 0, tc = 0;
    if (!bad) {
      auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
      const int reps = 3;
      double t0 = now();
      pool.run([&](int t) {
        __m256i acc = _mm256_setzero_si256();
        for (int r = 0; r < reps; r++)
          for (size_t i = 0; i < n; i += 32) acc = _mm256_add_epi64(acc, _mm256_load_si256((const __m256i*)(src[t] + i)));
        uint64_t o[4];
        _mm256_storeu_si256((__m256i*)o, acc);
        sink[t] = o[0] + o[1] + o[2] + o[3];
      });
      tr = (now() - t0) / reps;
      t0 = now();
      pool.run([&](int t) {
        for (int r = 0; r < reps; r++)
          for (size_t i = 0; i < n; i += 32)
            _mm256_stream_si256((__m256i*)(dst[t] + i), _mm256_load_si256((const __m256i*)(src[t] + i)));
        _mm_sfence();
      });
      tc = (now() - t0) / reps;
    }
    for (int t = 0; t < n_threads; t++) { free(src[t]); free(dst[t]); }
    if (bad) return MPTV_ERR_NOMEM;
    *read_gbs = (double)n * n_threads / tr / 1e9;
    *copy_gbs = (double)n * n_threads / tc / 1e9;  // payload; the traffic is twice that
    return sink[0] == 1 ? MPTV_OK : MPTV_OK;
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

const mptv_batch* mptv_host_batch_view(const mptv_host_batch* hb) { return hb ? &hb->view : nullptr; }
const uint8_t* mptv_host_batch_bad_root(const mptv_host_batch* hb) { return hb ? hb->bad_root : nullptr; }

void mptv_host_batch_free(mptv_host_batch* hb) {
  if (!hb) return;
  for (int i = 0; i < 8; i++) {
    if (!hb->blocks[i]) continue;
    if (hb->pinned) cudaFreeHost(hb->blocks[i]); else free(hb->blocks[i]);
  }
  delete hb;
}

uint32_t mptv_rlp_index(uint64_t index, uint8_t out[9]) {
  Out o{out, 9};
  put_u64(o, index);
  return (uint32_t)o.n;
}

uint64_t mptv_encode_receipt(int prefix, int status, uint64_t cumulative_gas_used, const uint8_t* bloom256,
                             const mptv_log* logs, uint32_t n_logs, uint8_t* out, uint64_t cap) {
  uint64_t logs_payload = 0;
  for (uint32_t i = 0; i < n_logs; i++) { const uint64_t p = log_payload(logs[i]); logs_payload += hdr_len(p) + p; }
  const uint64_t payload = 1 + u64_len(cumulative_gas_used) + 3 + 256 + hdr_len(logs_payload) + logs_payload;
  Out o{out, out ? cap : 0};
  if (prefix >= 0) o.put((uint8_t)prefix);
  put_hdr(o, payload, true);
  o.put(status ? 0x01 : 0x80);  // bool: alloy-rlp encodes false as the empty string
  put_u64(o, cumulative_gas_used);
  put_hdr(o, 256, false);
  o.put(bloom256, 256);
  put_hdr(o, logs_payload, true);
  for (uint32_t i = 0; i < n_logs; i++) {
    const mptv_log& l = logs[i];
    put_hdr(o, log_payload(l), true);
    o.put(0x94);
    o.put(l.address, 20);
    put_hdr(o, 33ull * l.n_topics, true);
    for (uint32_t t = 0; t < l.n_topics; t++) { o.put(0xa0); o.put(l.topics + 32 * t, 32); }
    put_str(o, l.data, l.data_len);
  }
  return o.n;
}

}  // extern "C"
