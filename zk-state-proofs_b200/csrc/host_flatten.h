// host_flatten.h -- internal: the single-pass, de-duplicating chunk builder of the host-fed entries.
//
// What crosses PCIe per proof is the bound of the host-fed path (one B200 hashes ~0.5 TB/s of
// node bytes, a PCIe 5 x16 link moves 55 GB/s), and most of it is redundant: every proof against one trie carries
// its own copy of the root node and of the full branches below it (config 2: 7.49 M supplied nodes, 2.08 M distinct;
// the duplicates are the 532-byte ones).  The builder therefore ALIASES a node that is byte-identical to one
// already placed in the same chunk: node_off[k] points at the first copy and nothing is written or copied for it.
// The device still hashes every supplied node (K1 runs over node indices; aliased reads hit L2), so the verdicts,
// the Keccak-f count W_perm and the headline semantics are untouched -- this is transfer de-duplication, not the
// "dedup_nodes" mode that shares digests.
//
// Identity is decided exactly: a sampled 64-bit fingerprint selects candidates in a shared lock-free table, a full
// memcmp against the first copy's source bytes (cache-hot: it was read moments ago) confirms.  The table holds a
// chunk's nodes only (epoch-tagged, so starting a chunk costs nothing) and stays cache resident.
//
// One pass over the input: phase 0 reads each blob's node count (4 bytes) to size the index arrays and to bound
// the bytes per worker; phase 1 streams every blob once -- fingerprint, probe, then either memcmp (duplicate) or a
// non-temporal copy into the worker's own region of the page-locked staging block.  Workers never share a cache
// line of the staging block and take no locks; the regions' used prefixes go to the device as separate copies.
#pragma once
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <utility>
#include <vector>

#include "../../include/mptv.h"
#include "host_codec.h"
#include "kernels.h"

namespace mptv {

#ifndef MPTV_STREAM_AHEAD
#define MPTV_STREAM_AHEAD 4096
#endif
constexpr uint32_t kStreamAhead = MPTV_STREAM_AHEAD;  // software prefetch distance over the blobs, bytes
constexpr uint32_t kDedupMinLen = 128;  // shorter nodes (leaves, 2-3 child branches) are mostly one-offs: not worth a probe

// ---- the candidate table: open addressing, insert-only within an epoch, shared by the workers of one chunk
struct DedupEntry {
  std::atomic<uint64_t> key;    // epoch << 48 | fingerprint >> 16     (epoch != current: the entry is free)
  std::atomic<uint64_t> ready;  // epoch << 48 | first copy's offset in 16-byte units, published after src / len
  const uint8_t* src;           // the first copy's bytes in the caller's input (what memcmp reads)
  uint32_t len, pad;
};
static_assert(sizeof(DedupEntry) == 32, "two entries per cache line");

class DedupTable {
 public:
  DedupTable() = default;
  DedupTable(const DedupTable&) = delete;
  DedupTable& operator=(const DedupTable&) = delete;
  DedupTable(DedupTable&& o) noexcept : tab_(o.tab_), mask_(o.mask_), epoch_(o.epoch_) { o.tab_ = nullptr; o.mask_ = 0; }
  ~DedupTable() { release(); }
  void release();
  bool reserve(size_t entries);  // power of two >= entries; false = out of memory (then lookups just miss)
  void new_epoch();              // every recorded node is forgotten (O(1) until the 16-bit epoch wraps)
  size_t capacity() const { return mask_ ? mask_ + 1 : 0; }
  // An identical node recorded in this epoch?  true: *off16 = its offset (16-byte units).  false: the node was
  // recorded with offset my_off16 (or the table is full / the probe budget ran out -- then it simply is not shared).
  bool find_or_insert(const uint8_t* p, uint32_t len, uint64_t fingerprint, uint32_t my_off16, uint32_t* off16);
  // bring the entry a fingerprint maps to into the cache ahead of find_or_insert
  void prefetch(uint64_t fingerprint) const {
    if (tab_) __builtin_prefetch(&tab_[(size_t)fingerprint & mask_]);
  }
  // second stage of the look-ahead: if the first entry of the probe sequence already holds this fingerprint, bring
  // the first copy's bytes (what memcmp will read) into the cache too.  A hint only: nothing is decided here.
  void prefetch_source(uint64_t fingerprint, uint32_t len) const {
    if (!tab_) return;
    const DedupEntry& e = tab_[(size_t)fingerprint & mask_];
    if (e.key.load(std::memory_order_relaxed) != ((epoch_ << 48) | (fingerprint >> 16))) return;
    if ((e.ready.load(std::memory_order_acquire) >> 48) != epoch_) return;
    const uint8_t* s = e.src;
    for (uint32_t o = 0; o < len; o += 64) __builtin_prefetch(s + o);
  }

 private:
  DedupEntry* tab_ = nullptr;
  size_t mask_ = 0;
  uint64_t epoch_ = 0;
};

// 64-bit fingerprint of a node of >= kDedupMinLen bytes from its length and 16 samples of 8 bytes spread evenly over
// it (one per child slot of a full branch, so versions of a node that differ in a single child land in different
// buckets).  Touches lines the caller reads anyway.
inline uint64_t node_fingerprint(const uint8_t* p, uint32_t len) {
  const uint32_t step = (len - 8) / 15;
  uint64_t h0 = 0x9E3779B97F4A7C15ull * len, h1 = 0xC2B2AE3D27D4EB4Full ^ len;
  for (int i = 0; i < 16; i += 2) {
    uint64_t a, b;
    memcpy(&a, p + (size_t)i * step, 8);
    memcpy(&b, p + (size_t)(i + 1) * step, 8);
    h0 = (((h0 << 23) | (h0 >> 41)) ^ a) * 0x9E3779B97F4A7C15ull;
    h1 = (((h1 << 29) | (h1 >> 35)) ^ b) * 0xC2B2AE3D27D4EB4Full;
  }
  uint64_t h = h0 ^ ((h1 << 32) | (h1 >> 32));
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  return h ^ (h >> 32);
}

// ---- layout of one flattened chunk inside a staging block (identical on the device)
struct ChunkLayout {
  uint64_t np = 0, nn = 0;
  size_t o_off = 0, o_len = 0, o_pf = 0, o_roots = 0, o_koff = 0, o_klen = 0;  // index arrays
  size_t index_end = 0;  // [0, index_end) is one contiguous copy
  size_t total = 0;      // extent of the block (worst case: nothing aliased)
  // pull mode (the input is page-locked: the device fetches the placed bytes itself): one gather record per placed
  // node / key in [o_gather, o_gather + 16 n_gather); the host block ends at host_total, the byte regions exist on
  // the device only
  size_t o_gather = 0, n_gather = 0, host_total = 0;
  // borsh(StorageProofInput) chunks (flatten_storage_chunk): the proofs form groups -- an account proof and the storage
  // proofs that take their root from its verified leaf -- and the storage keys are hashed on the device:
  // root_from_proof i32[np] (chunk-relative, -1 = none), hash_key u8[np]; device only: the key records the kernels read
  // after k_prepare_keys (offset u32[np], length u32[np]) and the 32-byte hashed keys
  bool groups = false;
  size_t o_rfp = 0, o_hk = 0, o_hkoff = 0, o_hklen = 0, o_hashed = 0;
  std::vector<size_t> region_begin, region_used;  // per worker: its byte region and how much of it was written
  // The byte regions may live in a host block of their own (write-combining staging): then get_block sets this to that
  // block's address MINUS index_end, so that region offsets stay offsets of the one device pack.  Null: the regions
  // follow the index arrays in the block get_block returned.
  uint8_t* region_base = nullptr;
  // statistics of the build
  uint64_t node_bytes_supplied = 0;  // sum of the padded lengths of all supplied nodes
  uint64_t node_bytes_placed = 0;    // ... of those actually written (the rest are aliases)
  uint64_t nodes_aliased = 0;
};

struct FlattenStats {
  uint64_t chunks = 0, nodes = 0, nodes_aliased = 0, node_bytes_supplied = 0, node_bytes_placed = 0, h2d_bytes = 0;
};

// Flatten blobs [cs, ce) (borsh(MerkleProofInput), /root/reference/crypto-ops/src/types.rs:4-9) into a staging block.
//   get_block(host_bytes, device_bytes) -> base pointer of a host block of at least host_bytes (called once, by
//   worker 0, between the phases; the device copy needs device_bytes; nullptr = allocation failed).
// Arrays inside the block: node_off u64[nn] (offsets from the block base), node_len u32[nn], proof_first u32[np+1],
// roots u8[32 np], key_off u32[np] + key_len u32[np] (keys live in the byte regions), then the workers' byte regions.
// node_src[k] (host only, may be null) = position of node k's bytes in `blobs`; bad_root[i] = root_hash.len() != 32.
// Returns MPTV_OK, MPTV_ERR_ARG (malformed blob) or MPTV_ERR_NOMEM.
struct BorshChunkJob {
  const uint8_t* blobs;
  const uint64_t* blob_off;
  uint64_t cs, ce;
  DedupTable* table;  // null = no de-duplication
  bool pull = false;  // write gather records instead of copying bytes (the caller's blobs are device-accessible)
  // key offsets in 16-byte units (keys start on 16-byte boundaries): a block of up to 64 GiB instead of 4 GiB.  The
  // device reads byte offsets, so only mptv_flatten_borsh_ex, which re-packs the keys anyway, asks for this.
  bool key_off_16 = false;
};
// what the device's gather kernel executes: len bytes at blobs + src -> block + 16 * dst16, padded with zeros to 16
struct GatherRec { uint64_t src; uint32_t dst16; uint32_t len; };
static_assert(sizeof(GatherRec) == 16, "one uint4 on the device");
constexpr uint32_t kGatherUnused = 0xffffffffu;
template <class GetBlock>
int flatten_borsh_chunk(WorkerPool& pool, const BorshChunkJob& job, GetBlock get_block, ChunkLayout& L,
                        std::vector<uint64_t>* node_src, std::vector<uint8_t>* bad_root);

// ------------------------------------------------------------------------------------------ implementation
inline size_t up16z(size_t x) { return (x + 15) & ~(size_t)15; }
inline size_t up64z(size_t x) { return (x + 63) & ~(size_t)63; }

inline uint32_t rd_u32(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);  // the wire format is little-endian, and so is every host this library runs on
  return v;
}

// one worker's cursor into its region of the staging block
struct RegionWriter {
  uint8_t* base;       // block base
  size_t at, end;      // next free byte / end of my region (offsets from base, `at` always a multiple of 16)
  DedupTable* table;
  uint64_t supplied = 0, placed = 0, aliased = 0;
  // place one node: returns its offset from the block base
  // fp: node_fingerprint(src, len), computed (and its table entry prefetched) one node ahead by the caller
  inline uint64_t put_node(const uint8_t* src, uint32_t len, uint64_t fp) {
    const size_t padded = up16z(len);
    supplied += padded;
    if (table && len >= kDedupMinLen) {
      uint32_t off16;
      if (table->find_or_insert(src, len, fp, (uint32_t)(at >> 4), &off16)) {
        aliased++;
        return (uint64_t)off16 << 4;
      }
    }
    const size_t o = at;
    put_bytes(base + o, src, len);
    at += padded;
    placed += padded;
    return o;
  }
  inline size_t put_key(const uint8_t* src, uint32_t len) {
    const size_t o = at;
    if (len) put_bytes(base + o, src, len);
    at += up16z(len);
    return o;
  }
  // Every byte is written once and next read by the DMA engine, so it should go out with non-temporal stores (no
  // read-for-ownership of the staging lines).  Without a table the node is streamed directly.  With a table that
  // does not work: the claim of an entry is a locked instruction, which drains the write-combining buffers, so the
  // partly filled last line of the previous node's copy goes out as a partial write (measured on the B200 host:
  // 86 -> 48 ns a node on one core once the stores no longer meet the locked instructions,
  // profiles/r02_flatten_micro.txt).  The bytes are therefore collected in a small cache-resident bounce buffer and
  // leave it as whole 64-byte lines, several KB at a time, with no locked instruction in between.
  static constexpr size_t kBounce = 8192;
  alignas(64) uint8_t bounce[kBounce];
  size_t bn = 0;        // bytes waiting in the bounce buffer
  size_t flushed = 0;   // region offset (from base) the first waiting byte belongs to; 64-byte aligned
  bool bounce_on = false;
  inline void flush_lines(bool all) {
    size_t lines = bn & ~(size_t)63;
    if (all && (bn & 63)) {  // the last, partial line: zero-fill (the region has the slack)
      memset(bounce + bn, 0, 64 - (bn & 63));
      lines = (bn + 63) & ~(size_t)63;
    }
    uint8_t* dst = base + flushed;
    for (size_t o = 0; o < lines; o += 32)
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + o), _mm256_load_si256(reinterpret_cast<const __m256i*>(bounce + o)));
    flushed += lines;
    const size_t rest = lines < bn ? bn - lines : 0;
    if (rest) memcpy(bounce, bounce + lines, rest);  // < 64 bytes
    bn = rest;
  }
  inline void append(const uint8_t* src, size_t len) {
    while (len) {
      const size_t c = std::min(len, kBounce - bn);
      memcpy(bounce + bn, src, c);
      bn += c; src += c; len -= c;
      if (bn == kBounce) flush_lines(false);
    }
  }
  GatherRec* rec = nullptr;   // pull mode: my gather list (next free record), else null
  const uint8_t* src_base = nullptr;
  inline void put_bytes(uint8_t* dst, const uint8_t* src, uint32_t len) {
    if (rec) {  // the device fetches the bytes from the caller's page-locked blobs: nothing is read or written here
      rec->src = (uint64_t)(src - src_base); rec->dst16 = (uint32_t)((size_t)(dst - base) >> 4); rec->len = len;
      rec++;
      return;
    }
    if (!table) { copy_node_stream(dst, src, len); return; }
    if (!bounce_on) { bounce_on = true; flushed = (size_t)(dst - base); }  // first use: dst is my region's (64-byte aligned) start
    append(src, len);
    static const uint8_t zeros[16] = {0};
    if (len & 15u) append(zeros, 16 - (len & 15u));
  }
  // everything still waiting goes out (call once, after the last put_*)
  inline void finish() {
    if (bounce_on) flush_lines(true);
  }
};

template <class GetBlock>
int flatten_borsh_chunk(WorkerPool& pool, const BorshChunkJob& job, GetBlock get_block, ChunkLayout& L,
                        std::vector<uint64_t>* node_src, std::vector<uint8_t>* bad_root) {
  const int T = pool.size();
  const uint64_t np = job.ce - job.cs;
  struct Tot { uint64_t nodes, bound; };
  std::vector<Tot> tot(T);
  std::vector<RegionWriter> wr(T);
  std::atomic<int> err(0);
  uint8_t* block = nullptr;
  L = ChunkLayout();
  L.np = np;
  L.region_begin.assign(T, 0);
  L.region_used.assign(T, 0);
  const uint64_t per = (np + T - 1) / T;
  pool.run([&](int t) {
    const uint64_t lo = std::min(np, per * t), hi = std::min(np, lo + per);
    // ---- phase 0: node counts (first word of each blob) -> index bases and a bound on my bytes
    Tot my = {0, 0};
    for (uint64_t i = lo; i < hi; i++) {
      const uint64_t b0 = job.blob_off[job.cs + i], b1 = job.blob_off[job.cs + i + 1];
      if (b1 < b0 || b1 - b0 < 12) { err.store(MPTV_ERR_ARG); continue; }
      const uint64_t n = rd_u32(job.blobs + b0);
      if (n > (b1 - b0 - 12) / 4) { err.store(MPTV_ERR_ARG); continue; }  // each node costs at least its length word
      my.nodes += n;
      my.bound += (b1 - b0) + 16 * (n + 1);  // every node and the key padded to 16
    }
    tot[t] = my;
    pool.barrier();
    if (t == 0 && !err.load()) {
      uint64_t nn = 0;
      for (int k = 0; k < T; k++) nn += tot[k].nodes;
      if (nn > 0xfffffff0ull) err.store(MPTV_ERR_ARG);
      else {
        size_t o = 0;
        auto take = [&](size_t bytes) { const size_t at = o; o += up64z(bytes); return at; };
        L.nn = nn;
        L.o_off = take(8 * nn); L.o_len = take(4 * nn); L.o_pf = take(4 * (np + 1)); L.o_roots = take(32 * np);
        L.o_koff = take(4 * np); L.o_klen = take(4 * np);
        L.index_end = o;
        if (job.pull) {  // one record per node and per key at most, each worker's list behind the one before
          L.o_gather = o;
          L.n_gather = nn + np;
          o += up64z(16 * (nn + np));
          L.host_total = o;
        }
        for (int k = 0; k < T; k++) { L.region_begin[k] = o; o += up64z(tot[k].bound) + 64; }
        L.total = o + 64;
        if (!job.pull) L.host_total = L.total;
        // node offsets are kept in 16-byte units in 32 bits (the table), key offsets as 32-bit bytes unless key_off_16
        if (L.total > (job.key_off_16 ? 0xfffffff00ull : 0xffffff00ull)) err.store(MPTV_ERR_ARG);
        else {
          block = get_block(L.host_total, L.total);
          if (!block) err.store(MPTV_ERR_NOMEM);
          else {
            if (node_src && node_src->size() < nn) node_src->resize(nn + nn / 8);
            if (bad_root && bad_root->size() < np) bad_root->resize(np + np / 8);
          }
        }
      }
    }
    pool.barrier();
    if (err.load()) return;
    // ---- phase 1: one streaming pass over my blobs
    uint64_t k = 0;
    for (int q = 0; q < t; q++) k += tot[q].nodes;
    const uint64_t k_end = k + tot[t].nodes;  // (a caller that rewrites the blobs during the call must not get past its share of the arrays)
    uint64_t* node_off = reinterpret_cast<uint64_t*>(block + L.o_off);
    uint32_t* node_len = reinterpret_cast<uint32_t*>(block + L.o_len);
    uint32_t* proof_first = reinterpret_cast<uint32_t*>(block + L.o_pf);
    uint32_t* key_off = reinterpret_cast<uint32_t*>(block + L.o_koff);
    uint32_t* key_len = reinterpret_cast<uint32_t*>(block + L.o_klen);
    uint8_t* roots = block + L.o_roots;
    RegionWriter w;  // on my stack: the cursor and counters change with every node (no false sharing)
    w.base = L.region_base ? L.region_base : block; w.at = L.region_begin[t]; w.end = w.at + up64z(tot[t].bound); w.table = job.table;
    GatherRec* rec_end = nullptr;
    if (job.pull) {
      uint64_t r0 = 0;
      for (int q = 0; q < t; q++) r0 += tot[q].nodes + (std::min(np, per * (q + 1)) - std::min(np, per * q));
      w.rec = reinterpret_cast<GatherRec*>(block + L.o_gather) + r0;
      rec_end = w.rec + tot[t].nodes + (hi - lo);
      w.src_base = job.blobs;
    }
    // fingerprint of the node whose length word is at q, computed one node ahead so that its table entry is in the
    // cache by the time the node is placed (0 when the node is not a candidate or does not fit the blob)
    // (two stages: the entry is prefetched two nodes ahead, the first copy's bytes one node ahead)
    auto look_ahead = [&](const uint8_t* q, const uint8_t* end) -> uint64_t {
      if (!job.table || end - q < 4) return 0;
      const uint32_t len = rd_u32(q);
      if (len < kDedupMinLen || (uint64_t)(end - q - 4) < len) return 0;
      const uint64_t fp = node_fingerprint(q + 4, len);
      job.table->prefetch(fp);
      return fp;
    };
    // the node after the one whose length word is at q (nullptr when it does not fit the blob)
    auto next_node = [&](const uint8_t* q, const uint8_t* end) -> const uint8_t* {
      if (end - q < 4) return nullptr;
      const uint32_t len = rd_u32(q);
      return (uint64_t)(end - q - 4) < len ? nullptr : q + 4 + len;
    };
    for (uint64_t i = lo; i < hi; i++) {
      const uint8_t* p = job.blobs + job.blob_off[job.cs + i];
      const uint8_t* end = job.blobs + job.blob_off[job.cs + i + 1];
      const uint32_t n = rd_u32(p);
      p += 4;
      proof_first[i] = (uint32_t)k;
      bool ok = n <= k_end - k;
      uint64_t fp = n ? look_ahead(p, end) : 0;  // of node j
      const uint8_t* p1 = n > 1 ? next_node(p, end) : nullptr;
      uint64_t fp1 = p1 ? look_ahead(p1, end) : 0;  // of node j + 1
      for (uint32_t j = 0; ok && j < n; j++) {
        if (end - p < 4) { ok = false; break; }
        const uint32_t len = rd_u32(p);
        if ((uint64_t)(end - p - 4) < len || len > kMaxNodeLen) { ok = false; break; }
        // keep the input stream a fixed distance ahead of the reads (the sampled fingerprints of the nodes ahead touch
        // lines the hardware prefetcher has not reached yet)
        for (uint32_t pf = 0; pf < len + 4; pf += 64) __builtin_prefetch(p + kStreamAhead + pf);
        const uint8_t* p2 = (p1 && j + 2 < n) ? next_node(p1, end) : nullptr;
        const uint64_t fp2 = p2 ? look_ahead(p2, end) : 0;  // of node j + 2: its entry starts travelling now
        if (fp1) job.table->prefetch_source(fp1, rd_u32(p1));  // node j + 1: its entry has arrived, fetch the first copy
        node_off[k] = w.put_node(p + 4, len, fp);
        fp = fp1; fp1 = fp2; p1 = p2;
        node_len[k] = len;
        if (node_src) (*node_src)[k] = (uint64_t)(p + 4 - job.blobs);
        k++;
        p += 4 + len;
      }
      uint32_t rl = 0, kl = 0;
      if (ok && (end - p < 4 || (uint64_t)(end - p - 4) < (rl = rd_u32(p)))) ok = false;
      if (ok) {
        if (rl == 32) memcpy(roots + 32 * i, p + 4, 32); else memset(roots + 32 * i, 0, 32);
        if (bad_root) (*bad_root)[i] = rl != 32;
        p += 4 + rl;
        if (end - p < 4 || (uint64_t)(end - p - 4) < (kl = rd_u32(p))) ok = false;
      }
      if (ok) {
        const size_t ko = w.put_key(p + 4, kl);
        key_off[i] = job.key_off_16 ? (uint32_t)(ko >> 4) : (uint32_t)ko;
        key_len[i] = kl;
        p += 4 + kl;
        if (p != end) ok = false;  // borsh::from_slice rejects trailing bytes
      }
      if (!ok) { err.store(MPTV_ERR_ARG); return; }
    }
    // a node count that lied (more nodes announced than present) was caught above; fewer is impossible (k is exact)
    if (job.pull) {
      for (; w.rec < rec_end; w.rec++) { w.rec->src = 0; w.rec->dst16 = 0; w.rec->len = kGatherUnused; }  // aliased nodes left these free
    } else {
      if (w.table) { static const uint8_t z16[16] = {0}; w.put_bytes(w.base + w.at, z16, 16); }  // K1 stages whole 16-byte chunks:
      else memset(w.base + w.at, 0, 16);                                                        // the bytes after my last node must be readable
    }
    w.at += 16;
    w.finish();
    _mm_sfence();  // non-temporal stores: the DMA engine (or another thread) reads them next
    wr[t] = w;
  });
  if (err.load()) return err.load();
  reinterpret_cast<uint32_t*>(block + L.o_pf)[np] = (uint32_t)L.nn;
  for (int t = 0; t < T; t++) {
    L.region_used[t] = wr[t].at - L.region_begin[t];
    L.node_bytes_supplied += wr[t].supplied; L.node_bytes_placed += wr[t].placed; L.nodes_aliased += wr[t].aliased;
  }
  return MPTV_OK;
}

// ------------------------------------------------------------------------------------------ borsh(StorageProofInput)
// /root/reference/crypto-ops/src/types.rs:11-19: account_proof Vec<Vec<u8>>, storage_proofs Vec<Vec<Vec<u8>>>,
// root_hash Vec<u8>, account_key Vec<u8>, storage_keys Vec<Vec<u8>>, address_keccak [u8; 32] -- the input of the risc0
// storage guest (circuits/risc0-storage-proof/.../storage-circuit/src/main.rs:6-31), which verifies the account
// proof under address_keccak and then zips storage_proofs with storage_keys (the shorter list decides, main.rs:18-21),
// verifying each under keccak(key) against the account's storage_root.  account_key is never read by the guest.
//
// One input = 1 + min(#storage_proofs, #storage_keys) proofs.  The storage keys lie BEHIND the proofs in the blob, so
// how many of the storage proofs count is only known at the end: a first pass over the length prefixes (one cache
// line per node) indexes every input, and the streaming pass then places exactly the nodes that count.
struct StorageIndex {
  std::vector<uint64_t> proof_first, node_first;  // [n + 1] prefix sums over the inputs: the guest's proofs / their nodes
  std::vector<uint32_t> root_at;                  // [n] where root_hash's length word sits inside blob i
  uint64_t indexed = 0;                           // inputs [0, indexed) are indexed
  void reset(uint64_t n) { proof_first.assign(n + 1, 0); node_first.assign(n + 1, 0); root_at.assign(n, 0); indexed = 0; }
};

// One input's walk over its length prefixes as a resumable cursor: step() consumes ONE prefix (a node's, a key's, a
// count) and stops, so that a thread can keep several inputs in flight -- every prefix of a node sits on a cache line
// of its own, 4 + len bytes behind the one before, and a single input walked on its own is one dependent miss after
// another (measured on the 16-core host: 80 ms for the 27 M nodes of config 3 with one cursor per thread).
struct StorageCursor {
  enum Phase : uint8_t { kAccCount, kAccNode, kStoCount, kProofCount, kProofNode, kRoot, kAccountKey, kKeyCount, kKey, kDone, kBad };
  const uint8_t* start = nullptr;
  const uint8_t* p = nullptr;
  const uint8_t* end = nullptr;
  uint64_t input = 0;       // which input this is
  uint64_t rem = 0, m = 0, j = 0, k = 0, n_acc = 0;
  uint32_t root_at = 0;
  Phase phase = kDone;
  std::vector<uint64_t> cum;  // cum[j] = nodes of the first j storage proofs
  void begin(uint64_t i, const uint8_t* s, const uint8_t* e) {
    input = i; start = p = s; end = e; rem = m = j = k = n_acc = 0; root_at = 0;
    cum.clear(); cum.push_back(0);
    phase = (uint64_t)(e - s) > 0xfffffff0ull ? kBad : kAccCount;  // positions inside a blob are kept in 32 bits
  }
  bool active() const { return phase != kDone && phase != kBad; }
  // a count of elements that each cost at least 4 bytes
  bool count(uint64_t& out) {
    if (end - p < 4) return false;
    out = rd_u32(p);
    p += 4;
    return out <= (uint64_t)(end - p) / 4;
  }
  bool skip_bytes(bool node) {
    if (end - p < 4) return false;
    const uint32_t len = rd_u32(p);
    if ((uint64_t)(end - p - 4) < len || (node && len > kMaxNodeLen)) return false;
    p += 4 + len;
    return true;
  }
  void next_proof() { phase = j == m ? kRoot : kProofCount; }
  void step() {
    bool ok = true;
    switch (phase) {
      case kAccCount: ok = count(n_acc); rem = n_acc; phase = rem ? kAccNode : kStoCount; break;
      case kAccNode: ok = skip_bytes(true); if (--rem == 0) phase = kStoCount; break;
      case kStoCount: ok = count(m); j = 0; next_proof(); break;
      case kProofCount:
        ok = count(rem);
        cum.push_back(cum.back() + rem);
        if (rem) phase = kProofNode; else { j++; next_proof(); }
        break;
      case kProofNode: ok = skip_bytes(true); if (--rem == 0) { j++; next_proof(); } break;
      case kRoot: root_at = (uint32_t)(p - start); ok = skip_bytes(false); phase = kAccountKey; break;
      case kAccountKey: ok = skip_bytes(false); phase = kKeyCount; break;
      case kKeyCount: ok = count(k); rem = k; phase = rem ? kKey : kDone; break;
      case kKey: ok = skip_bytes(false); if (--rem == 0) phase = kDone; break;
      default: break;
    }
    if (!ok) phase = kBad;
    else if (phase == kDone && end - p != 32) phase = kBad;  // address_keccak, and nothing after it
  }
};

// index inputs [i0, i1) of an index sized for n inputs (StorageIndex::reset) whose entries below i0 are complete:
// MPTV_OK or MPTV_ERR_ARG (a blob is not what borsh::from_slice::<StorageProofInput> accepts).  Ranges must be indexed
// in ascending order (the prefix sums are extended from i0).
inline int skim_storage_range(WorkerPool& pool, const uint8_t* blobs, const uint64_t* blob_off, uint64_t i0, uint64_t i1, StorageIndex& idx) {
  const int T = pool.size();
  std::atomic<int> err(0);
  // equal shares of the BYTES (inputs differ in size by the number of storage proofs they carry)
  const uint64_t total = blob_off[i1] - blob_off[i0];
  pool.run([&](int t) {
    const uint64_t lo = (uint64_t)(std::lower_bound(blob_off + i0, blob_off + i1, blob_off[i0] + total / T * t) - blob_off);
    const uint64_t hi = t == T - 1 ? i1 : (uint64_t)(std::lower_bound(blob_off + i0, blob_off + i1, blob_off[i0] + total / T * (t + 1)) - blob_off);
    constexpr int kInFlight = 24;  // cursors per thread; their next prefixes are prefetched into L2 (more requests in flight than L1's fill buffers hold)
    StorageCursor cur[kInFlight];
    uint64_t next = lo;
    int live = 0;
    auto refill = [&](StorageCursor& c) {
      if (next >= hi) return false;
      if (blob_off[next + 1] < blob_off[next]) { err.store(MPTV_ERR_ARG); return false; }
      c.begin(next, blobs + blob_off[next], blobs + blob_off[next + 1]);
      next++;
      __builtin_prefetch(c.p, 0, 2);
      return true;
    };
    for (int c = 0; c < kInFlight; c++) live += refill(cur[c]) ? 1 : 0;
    while (live > 0 && !err.load(std::memory_order_relaxed)) {
      for (int ci = 0; ci < kInFlight; ci++) {
        StorageCursor& c = cur[ci];
        if (c.phase == StorageCursor::kDone && c.start == nullptr) continue;  // retired slot
        if (c.active()) {
          c.step();
          if (c.active()) { if (c.end - c.p >= 4) __builtin_prefetch(c.p, 0, 2); continue; }
        }
        if (c.phase == StorageCursor::kBad) { err.store(MPTV_ERR_ARG); return; }
        // finished: what the guest goes on to verify (main.rs:18-21: zip, the shorter list decides)
        const uint64_t used = std::min(c.m, c.k);
        idx.proof_first[c.input + 1] = 1 + used;
        idx.node_first[c.input + 1] = c.n_acc + c.cum[used];
        idx.root_at[c.input] = c.root_at;
        if (!refill(c)) { c.start = nullptr; c.phase = StorageCursor::kDone; live--; }
      }
    }
  });
  if (err.load()) return err.load();
  for (uint64_t i = i0; i < i1; i++) { idx.proof_first[i + 1] += idx.proof_first[i]; idx.node_first[i + 1] += idx.node_first[i]; }
  idx.indexed = i1;
  return MPTV_OK;
}

// index inputs [0, n)
inline int skim_storage_inputs(WorkerPool& pool, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, StorageIndex& idx) {
  idx.reset(n);
  return n ? skim_storage_range(pool, blobs, blob_off, 0, n, idx) : MPTV_OK;
}

struct StorageChunkJob {
  const uint8_t* blobs;
  const uint64_t* blob_off;
  uint64_t cs, ce;           // inputs of the chunk
  DedupTable* table;
  const StorageIndex* idx;   // of the whole call
  bool key_off_16 = false;   // as BorshChunkJob::key_off_16
};

// Flatten inputs [cs, ce) into a staging block laid out like flatten_borsh_chunk's, plus the group arrays (ChunkLayout).
// Proof order inside the chunk = the guest's: each input's account proof, then its storage proofs.
template <class GetBlock>
int flatten_storage_chunk(WorkerPool& pool, const StorageChunkJob& job, GetBlock get_block, ChunkLayout& L,
                          std::vector<uint64_t>* node_src, std::vector<uint8_t>* bad_root) {
  const int T = pool.size();
  const StorageIndex& X = *job.idx;
  const uint64_t ni = job.ce - job.cs;
  const uint64_t np = X.proof_first[job.ce] - X.proof_first[job.cs], nn = X.node_first[job.ce] - X.node_first[job.cs];
  std::vector<RegionWriter> wr(T);
  std::vector<uint64_t> bound(T, 0);
  std::atomic<int> err(0);
  uint8_t* block = nullptr;
  L = ChunkLayout();
  L.np = np; L.nn = nn; L.groups = true;
  L.region_begin.assign(T, 0);
  L.region_used.assign(T, 0);
  if (nn > 0xfffffff0ull || np > 0x7ffffff0ull) return MPTV_ERR_ARG;
  const uint64_t per = (ni + T - 1) / T;
  {
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += up64z(bytes); return at; };
    L.o_off = take(8 * nn); L.o_len = take(4 * nn); L.o_pf = take(4 * (np + 1)); L.o_roots = take(32 * np);
    L.o_koff = take(4 * np); L.o_klen = take(4 * np); L.o_rfp = take(4 * np); L.o_hk = take(np);
    L.index_end = o;
    L.o_hkoff = take(4 * np); L.o_hklen = take(4 * np); L.o_hashed = take(32 * np);  // written by the device
    for (int t = 0; t < T; t++) {
      const uint64_t lo = job.cs + std::min(ni, per * t), hi = job.cs + std::min(ni, per * (t + 1));
      // every node and every key padded to 16: the blob bytes bound what is placed
      bound[t] = (job.blob_off[hi] - job.blob_off[lo]) + 16 * ((X.node_first[hi] - X.node_first[lo]) + (X.proof_first[hi] - X.proof_first[lo]) + 1);
      L.region_begin[t] = o;
      o += up64z(bound[t]) + 64;
    }
    L.total = o + 64;
    L.host_total = L.total;
    if (L.total > (job.key_off_16 ? 0xfffffff00ull : 0xffffff00ull)) return MPTV_ERR_ARG;  // as flatten_borsh_chunk
    block = get_block(L.host_total, L.total);
    if (!block) return MPTV_ERR_NOMEM;
    if (node_src && node_src->size() < nn) node_src->resize(nn + nn / 8);
    if (bad_root && bad_root->size() < np) bad_root->resize(np + np / 8);
  }
  pool.run([&](int t) {
    const uint64_t lo = job.cs + std::min(ni, per * t), hi = job.cs + std::min(ni, per * (t + 1));
    uint64_t k = X.node_first[lo] - X.node_first[job.cs];     // next node index of the chunk
    uint64_t pi = X.proof_first[lo] - X.proof_first[job.cs];  // next proof index of the chunk
    // (a caller that rewrites the blobs between the index pass and this one must not get past my share of the arrays)
    const uint64_t k_end = X.node_first[hi] - X.node_first[job.cs], pi_end = X.proof_first[hi] - X.proof_first[job.cs];
    uint64_t* node_off = reinterpret_cast<uint64_t*>(block + L.o_off);
    uint32_t* node_len = reinterpret_cast<uint32_t*>(block + L.o_len);
    uint32_t* proof_first = reinterpret_cast<uint32_t*>(block + L.o_pf);
    uint32_t* key_off = reinterpret_cast<uint32_t*>(block + L.o_koff);
    uint32_t* key_len = reinterpret_cast<uint32_t*>(block + L.o_klen);
    int32_t* rfp = reinterpret_cast<int32_t*>(block + L.o_rfp);
    uint8_t* hk = block + L.o_hk;
    uint8_t* roots = block + L.o_roots;
    RegionWriter w;
    w.base = L.region_base ? L.region_base : block; w.at = L.region_begin[t]; w.end = w.at + up64z(bound[t]); w.table = job.table;
    auto look_ahead = [&](const uint8_t* q, const uint8_t* end) -> uint64_t {
      if (!job.table || end - q < 4) return 0;
      const uint32_t len = rd_u32(q);
      if (len < kDedupMinLen || (uint64_t)(end - q - 4) < len) return 0;
      const uint64_t fp = node_fingerprint(q + 4, len);
      job.table->prefetch(fp);
      return fp;
    };
    auto next_node = [&](const uint8_t* q, const uint8_t* end) -> const uint8_t* {
      if (end - q < 4) return nullptr;
      const uint32_t len = rd_u32(q);
      return (uint64_t)(end - q - 4) < len ? nullptr : q + 4 + len;
    };
    // the nodes of the Vec<Vec<u8>> at p become the next proof of the chunk (same look-ahead as flatten_borsh_chunk)
    auto place_proof = [&](const uint8_t*& p, const uint8_t* end) -> bool {
      if (end - p < 4 || pi >= pi_end) return false;
      const uint32_t n = rd_u32(p);
      p += 4;
      if (n > k_end - k) return false;
      proof_first[pi] = (uint32_t)k;
      uint64_t fp = n ? look_ahead(p, end) : 0;
      const uint8_t* p1 = n > 1 ? next_node(p, end) : nullptr;
      uint64_t fp1 = p1 ? look_ahead(p1, end) : 0;
      for (uint32_t j = 0; j < n; j++) {
        if (end - p < 4) return false;
        const uint32_t len = rd_u32(p);
        if ((uint64_t)(end - p - 4) < len || len > kMaxNodeLen) return false;
        for (uint32_t pf = 0; pf < len + 4; pf += 64) __builtin_prefetch(p + kStreamAhead + pf);
        const uint8_t* p2 = (p1 && j + 2 < n) ? next_node(p1, end) : nullptr;
        const uint64_t fp2 = p2 ? look_ahead(p2, end) : 0;
        if (fp1) job.table->prefetch_source(fp1, rd_u32(p1));
        node_off[k] = w.put_node(p + 4, len, fp);
        fp = fp1; fp1 = fp2; p1 = p2;
        node_len[k] = len;
        if (node_src) (*node_src)[k] = (uint64_t)(p + 4 - job.blobs);
        k++;
        p += 4 + len;
      }
      pi++;
      return true;
    };
    for (uint64_t i = lo; i < hi; i++) {
      const uint8_t* start = job.blobs + job.blob_off[i];
      const uint8_t* end = job.blobs + job.blob_off[i + 1];
      const uint8_t* p = start;
      const uint64_t acc = pi, used = X.proof_first[i + 1] - X.proof_first[i] - 1;
      bool ok = place_proof(p, end);  // account_proof
      if (ok && end - p < 4) ok = false;
      if (ok) p += 4;                 // storage_proofs.len(): the index knows how many count
      for (uint64_t j = 0; ok && j < used; j++) ok = place_proof(p, end);
      if (ok && X.root_at[i] > (uint64_t)(end - start) - 4) ok = false;
      if (!ok) { err.store(MPTV_ERR_ARG); return; }
      p = start + X.root_at[i];  // (past the storage proofs that have no key)
      const uint32_t rl = rd_u32(p);
      if ((uint64_t)(end - p - 4) < rl) { err.store(MPTV_ERR_ARG); return; }
      if (rl == 32) memcpy(roots + 32 * acc, p + 4, 32); else memset(roots + 32 * acc, 0, 32);
      if (bad_root) (*bad_root)[acc] = rl != 32;  // the guest's try_into().unwrap() (main.rs:11)
      p += 4 + rl;
      uint32_t al = 0;
      if (end - p < 4 || (uint64_t)(end - p - 4) < (al = rd_u32(p))) { err.store(MPTV_ERR_ARG); return; }
      p += 4 + al;               // account_key: never read by the guest
      if (end - p < 4) { err.store(MPTV_ERR_ARG); return; }
      p += 4;                    // storage_keys.len()
      rfp[acc] = -1; hk[acc] = 0;
      for (uint64_t j = 0; j < used; j++) {
        uint32_t kl = 0;
        if (end - p < 4 || (uint64_t)(end - p - 4) < (kl = rd_u32(p))) { err.store(MPTV_ERR_ARG); return; }
        const uint64_t q = acc + 1 + j;
        const size_t ko = w.put_key(p + 4, kl);  // the RAW slot: digest_keccak(&key) runs on the device (main.rs:26)
        key_off[q] = job.key_off_16 ? (uint32_t)(ko >> 4) : (uint32_t)ko;
        key_len[q] = kl;
        memset(roots + 32 * q, 0, 32);                // taken on the device from the verified account leaf
        rfp[q] = (int32_t)acc; hk[q] = 1;
        if (bad_root) (*bad_root)[q] = 0;
        p += 4 + kl;
      }
      if (end - start < 32) { err.store(MPTV_ERR_ARG); return; }
      const size_t ako = w.put_key(end - 32, 32);  // address_keccak
      key_off[acc] = job.key_off_16 ? (uint32_t)(ako >> 4) : (uint32_t)ako;
      key_len[acc] = 32;
    }
    if (w.table) { static const uint8_t z16[16] = {0}; w.put_bytes(w.base + w.at, z16, 16); }
    else memset(w.base + w.at, 0, 16);
    w.at += 16;
    w.finish();
    _mm_sfence();
    wr[t] = w;
  });
  if (err.load()) return err.load();
  reinterpret_cast<uint32_t*>(block + L.o_pf)[np] = (uint32_t)nn;
  for (int t = 0; t < T; t++) {
    L.region_used[t] = wr[t].at - L.region_begin[t];
    L.node_bytes_supplied += wr[t].supplied; L.node_bytes_placed += wr[t].placed; L.nodes_aliased += wr[t].aliased;
  }
  return MPTV_OK;
}

}  // namespace mptv
