// rebuild_kernels.cu -- K4: level-synchronous Merkle-Patricia-Trie rebuild (root hashes, node arena).
//
// Replaces, for a whole batch of tries at once, what trie-utils does per block with eth_trie:
//     let mut trie = EthTrie::new(memdb);
//     for (i, tx) in txs { trie.insert(&rlp(i), &tx.eip2718_encoded()) }      transaction.rs:41-63
//     trie.root_hash()                                                         transaction.rs:66
//   /root/reference/trie-utils/src/proofs/transaction.rs:41-68, proofs/receipt.rs:49-86
//
// The MPT of a key/value set is unique, so the reference's sequential node surgery (insert_at:
// leaf split, extension split, branch descent) is not replayed.  Instead:
//   k_trie_structure  one CTA per trie: bitonic sort of the items by (key, insertion index),
//                     last-write-wins + delete filter (insert(k, b"") removes k), adjacent-key LCP
//                     in nibbles, then the trie skeleton in BFS order -- a branch per LCP interval,
//                     an extension where a branch sits more than one nibble below its parent, a
//                     leaf per key, a key that ends at a branch is that branch's value -- and, in
//                     reverse BFS order, every node's exact encoded length and height.
//   k_trie_level_*    counting sort of all nodes of the batch by (height, hashed?)
//   k_trie_encode     per level, bottom-up: one warp per node writes its RLP encoding into the node
//                     arena (children < 32 bytes embedded, others referenced by the digest that the
//                     previous level's Keccak launch produced; commit()/write_node semantics)
//   K0 + K1           (keccak_kernels.cu) hash all >= 32-byte nodes of the level in one launch
// The leaf level carries > 95 % of the bytes and permutations of a tx / receipt trie.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"
#include "trie_rec.cuh"

namespace mptv {

// ------------------------------------------------------------------ input scan
__global__ void __launch_bounds__(256) k_trie_scan_input(const TrieBatchDev in, TrieSummary* sum) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mk = 0, mn = 0;
  if (i < in.n_items) mk = in.key_off[i + 1] - in.key_off[i];
  if (i < in.n_tries) mn = in.trie_first[i + 1] - in.trie_first[i];
  for (int o = 16; o; o >>= 1) {
    mk = max(mk, __shfl_xor_sync(0xffffffffu, mk, o));
    mn = max(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mk) atomicMax(&sum->max_key_len, mk);
    if (mn) atomicMax(&sum->max_items, mn);
  }
}

// ------------------------------------------------------------------ structure
namespace {

struct SortView {
  uint64_t* pre;  // first 8 key bytes, big-endian, zero padded
  uint32_t* idx;  // local item index (insertion order); kNoItem = padding, sorts last
  uint8_t* kl;    // key length in bytes
};

// strict "a sorts after b" over (key bytes, key length, insertion index)
__device__ bool sort_greater(const SortView& v, uint32_t a, uint32_t b, const uint8_t* key_bytes,
                             const uint32_t* key_off /* of the trie's first item */) {
  const uint32_t ia = v.idx[a], ib = v.idx[b];
  if (ia == kNoItem || ib == kNoItem) return ia == kNoItem && ib != kNoItem;
  const uint64_t pa = v.pre[a], pb = v.pre[b];
  if (pa != pb) return pa > pb;
  const uint32_t la = v.kl[a], lb = v.kl[b];
  if (la > 8 && lb > 8) {
    const uint8_t* ka = key_bytes + key_off[ia];
    const uint8_t* kb = key_bytes + key_off[ib];
    const uint32_t n = la < lb ? la : lb;
    for (uint32_t i = 8; i < n; i++) {
      const uint32_t x = ka[i], y = kb[i];
      if (x != y) return x > y;
    }
  }
  if (la != lb) return la > lb;  // equal up to the shorter key: the prefix sorts first
  return ia > ib;
}

__device__ bool keys_equal(const SortView& v, uint32_t a, uint32_t b, const uint8_t* key_bytes,
                           const uint32_t* key_off) {
  if (v.pre[a] != v.pre[b] || v.kl[a] != v.kl[b]) return false;
  const uint32_t l = v.kl[a];
  const uint8_t* ka = key_bytes + key_off[v.idx[a]];
  const uint8_t* kb = key_bytes + key_off[v.idx[b]];
  for (uint32_t i = 8; i < l; i++)
    if (ka[i] != kb[i]) return false;
  return true;
}

// nibble i of sorted key j
__device__ __forceinline__ uint32_t key_nib(const SortView& v, uint32_t j, uint32_t i, const uint8_t* key_bytes,
                                            const uint32_t* key_off) {
  const uint32_t by = i >> 1;
  const uint32_t b = by < 8 ? (uint32_t)(v.pre[j] >> (56 - 8 * by)) & 0xffu : key_bytes[key_off[v.idx[j]] + by];
  return (i & 1) ? (b & 15u) : (b >> 4);
}

// common prefix of sorted keys j and j+1 in nibbles
__device__ uint32_t lcp_nibbles(const SortView& v, uint32_t j, const uint8_t* key_bytes, const uint32_t* key_off) {
  const uint32_t la = v.kl[j], lb = v.kl[j + 1];
  const uint32_t cap = 2 * (la < lb ? la : lb);
  const uint64_t x = v.pre[j] ^ v.pre[j + 1];
  uint32_t c;
  if (x) c = (uint32_t)__clzll((long long)x) >> 2;
  else {
    c = 16;
    if (la > 8 && lb > 8) {
      const uint8_t* ka = key_bytes + key_off[v.idx[j]];
      const uint8_t* kb = key_bytes + key_off[v.idx[j + 1]];
      const uint32_t n = la < lb ? la : lb;
      uint32_t i = 8;
      while (i < n && ka[i] == kb[i]) i++;
      c = 2 * i;
      if (i < n && ((ka[i] ^ kb[i]) >> 4) == 0) c++;
    }
  }
  return c < cap ? c : cap;
}

}  // namespace

// Shared-memory plan: [pre u64 | idx u32 | kl u8 | lcp u8] x cap for the sort, then (node_cap > 0) the
// node records and lengths of this trie, so that the two serial passes of thread 0 (skeleton, inner-node
// sizes) never wait on global memory; tries too large for that (3 n nodes x 20 B) work out of the global
// node table through the same pointers.
__global__ void __launch_bounds__(kTrieThreads) k_trie_structure(const TrieBatchDev in, const TrieWork w, uint32_t cap,
                                                                 uint32_t node_cap) {
  extern __shared__ __align__(16) uint8_t tsm[];
  SortView v;
  v.pre = reinterpret_cast<uint64_t*>(tsm);
  v.idx = reinterpret_cast<uint32_t*>(tsm + 8ull * cap);
  v.kl = tsm + 12ull * cap;
  uint8_t* lcp = tsm + 13ull * cap;
  __shared__ uint32_t s_hist[2 * kMaxLevels];
  __shared__ uint32_t s_scan[kTrieThreads];
  __shared__ uint32_t s_m;
  __shared__ unsigned long long s_arena_base, s_perms, s_hashed;

  const uint32_t t = blockIdx.x;
  const uint32_t first = in.trie_first[t];
  const uint32_t n = in.trie_first[t + 1] - first;
  const uint32_t tid = threadIdx.x;
  const uint32_t* koff = in.key_off + first;  // local item i: key_bytes[koff[i] .. koff[i+1])
  const uint32_t base = 3u * first;           // node ids of this trie: base + BFS index
  for (uint32_t i = tid; i < 2 * kMaxLevels; i += kTrieThreads) s_hist[i] = 0;
  if (tid == 0) { s_perms = 0; s_hashed = 0; }
  if (n == 0) {
    if (tid == 0) w.tcount[t] = 0;
    return;
  }
  const bool in_smem = 3u * n <= node_cap;
  uint4* R = in_smem ? reinterpret_cast<uint4*>(tsm + ((14ull * cap + 15) & ~15ull)) : w.rec + base;
  uint32_t* Ln = in_smem ? reinterpret_cast<uint32_t*>(tsm + ((14ull * cap + 15) & ~15ull) + 16ull * node_cap) : w.len + base;
  uint32_t np2 = 2;
  while (np2 < n) np2 <<= 1;

  // ---- load: key prefixes, lengths, identity permutation; and is item i's key exactly rlp(i)?
  // (that is how trie-utils keys every transaction / receipt trie: alloy_rlp::encode(index) in block
  // order, transaction.rs:45) -- then the sorted order is known without sorting
  bool indexed = true;
  for (uint32_t i = tid; i < np2; i += kTrieThreads) {
    uint64_t p = ~0ull;
    uint32_t id = kNoItem, l = 255;
    if (i < n) {
      id = i;
      l = koff[i + 1] - koff[i];
      p = 0;
      const uint8_t* k = in.key_bytes + koff[i];
      for (uint32_t b = 0; b < 8 && b < l; b++) p |= (uint64_t)k[b] << (56 - 8 * b);
      // alloy_rlp::encode(i) for i < 65536: 0x80 | i (< 0x80) | 0x81 i | 0x82 hi lo
      const uint64_t want = i == 0 ? 0x80ull << 56
                            : i < 0x80 ? (uint64_t)i << 56
                            : i < 0x100 ? (0x81ull << 56) | ((uint64_t)i << 48)
                                        : (0x82ull << 56) | ((uint64_t)i << 40);
      const uint32_t wl = i < 0x80 ? 1u : (i < 0x100 ? 2u : 3u);
      indexed = indexed && p == want && l == wl;
    }
    v.pre[i] = p; v.idx[i] = id; v.kl[i] = (uint8_t)l;
  }
  indexed = __syncthreads_and(indexed);

  if (indexed) {
    // byte order of rlp(i): 0x01..0x7f (items 1..127), 0x80 (item 0), 0x81 0x80.. (128..255), 0x82.. (256..):
    // the first k = min(n, 128) entries rotate left by one, everything else is already in place
    static_assert(kTrieThreads >= 128, "the rotation below gives one of the first 128 entries to each thread");
    const uint32_t k = n < 128 ? n : 128;
    uint64_t p = 0; uint32_t q = 0, l = 0;
    if (tid < k) { const uint32_t src = tid + 1 == k ? 0 : tid + 1; p = v.pre[src]; q = v.idx[src]; l = v.kl[src]; }
    __syncthreads();
    if (tid < k) { v.pre[tid] = p; v.idx[tid] = q; v.kl[tid] = (uint8_t)l; }
    __syncthreads();
  } else {
    // ---- bitonic sort by (key, insertion index): one compare-exchange per thread per step
    for (uint32_t k = 2; k <= np2; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t t2 = tid; t2 < np2 / 2; t2 += kTrieThreads) {
          const uint32_t i = ((t2 & ~(j - 1)) << 1) | (t2 & (j - 1)), x = i | j;
          const bool up = (i & k) == 0;
          if (sort_greater(v, i, x, in.key_bytes, koff) == up) {
            const uint64_t p = v.pre[i]; v.pre[i] = v.pre[x]; v.pre[x] = p;
            const uint32_t q = v.idx[i]; v.idx[i] = v.idx[x]; v.idx[x] = q;
            const uint8_t l = v.kl[i]; v.kl[i] = v.kl[x]; v.kl[x] = l;
          }
        }
        __syncthreads();
      }
    }
  }

  // ---- last write wins, empty value deletes; compact the live keys in place
  {
    const uint32_t seg = (n + kTrieThreads - 1) / kTrieThreads;  // <= kTrieMaxItems / kTrieThreads
    const uint32_t s0 = tid * seg, s1 = min(n, s0 + seg);
    uint64_t kp[kTrieMaxItems / kTrieThreads];
    uint32_t ki[kTrieMaxItems / kTrieThreads];
    uint8_t kn[kTrieMaxItems / kTrieThreads];
    uint32_t cnt = 0;
    for (uint32_t s = s0; s < s1; s++) {
      const bool last = (s + 1 == n) || !keys_equal(v, s, s + 1, in.key_bytes, koff);
      if (last && in.value_len[first + v.idx[s]] != 0) {
        kp[cnt] = v.pre[s]; ki[cnt] = v.idx[s]; kn[cnt] = v.kl[s];
        cnt++;
      }
    }
    s_scan[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (uint32_t i = 0; i < kTrieThreads; i++) { const uint32_t c = s_scan[i]; s_scan[i] = acc; acc += c; }
      s_m = acc;
    }
    __syncthreads();
    const uint32_t o = s_scan[tid];
    for (uint32_t c = 0; c < cnt; c++) { v.pre[o + c] = kp[c]; v.idx[o + c] = ki[c]; v.kl[o + c] = kn[c]; }
    __syncthreads();
  }
  const uint32_t m = s_m;
  for (uint32_t j = tid; j + 1 < m; j += kTrieThreads) lcp[j] = (uint8_t)lcp_nibbles(v, j, in.key_bytes, koff);
  __syncthreads();

  // ---- skeleton: level-synchronous BFS over key ranges inside the CTA.  A pending entry sits in R[id]
  //      as x = lo | hi << 16, y = depth until it is expanded into the node's record.  Each round
  //      expands the current frontier [head, tail): leaves one per thread, inner nodes one per warp
  //      (the lanes scan the range's LCP values together: min-reduce for the extension test, ballot +
  //      popcount to cut the range into the branch's child groups, which get consecutive ids).
  __shared__ uint32_t s_tail;
  __shared__ uint32_t s_round[kMaxLevels + 2];  // s_round[r] = first node id created for round r
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t kWarps = kTrieThreads / 32;
  if (tid == 0) {
    s_tail = m ? 1u : 0u;
    if (m) R[0] = make_uint4(m << 16, 0, 0, kPending);
    s_round[0] = 0;
  }
  __syncthreads();
  uint32_t rounds = 0;
  for (uint32_t head = 0;;) {
    const uint32_t tail = s_tail;
    __syncthreads();  // everyone has read the frontier's end before anyone appends to it
    if (head == tail) break;
    if (tid == 0) s_round[rounds + 1] = tail;
    for (uint32_t e = head + tid; e < tail; e += kTrieThreads) {
      const uint4 q = R[e];
      const uint32_t lo = q.x & 0xffffu, hi = q.x >> 16, d = q.y;
      if (hi - lo == 1)
        R[e] = make_uint4(kTLeaf | (d << 16) | ((2u * v.kl[lo] - d) << 24), first + v.idx[lo], 0, 0);
    }
    __syncthreads();  // leaf records are in place; what is still pending is an inner node
    for (uint32_t e = head + warp; e < tail; e += kWarps) {
      const uint4 q = R[e];
      if (q.w != kPending) continue;
      const uint32_t lo = q.x & 0xffffu, hi = q.x >> 16, d = q.y;
      uint32_t c = 255;
      for (uint32_t j = lo + lane; j + 1 < hi; j += 32) c = min(c, (uint32_t)lcp[j]);
      for (int o = 16; o; o >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
      uint4 r = make_uint4(0, kNoItem, 0, 0);
      if (c > d) {
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(&s_tail, 1u);
        at = __shfl_sync(0xffffffffu, at, 0);
        r.x = kTExt | (d << 16) | ((c - d) << 24);
        r.y = first + v.idx[lo];
        r.z = base + at;
        if (lane == 0) R[at] = make_uint4(lo | (hi << 16), c, 0, kPending);
      } else {
        uint32_t l2 = lo;
        if (2u * v.kl[lo] == d) { r.y = first + v.idx[lo]; l2 = lo + 1; }  // this key ends here: branch value
        // pass 1: number of child groups = boundaries j in [l2, hi) with j + 1 == hi or lcp[j] == d
        uint32_t nc = 0;
        for (uint32_t b0 = l2; b0 < hi; b0 += 32) {
          const uint32_t j = b0 + lane;
          const bool bd = j < hi && (j + 1 == hi || lcp[j] == d);
          nc += __popc(__ballot_sync(0xffffffffu, bd));
        }
        uint32_t fc = 0;
        if (lane == 0) fc = atomicAdd(&s_tail, nc);
        fc = __shfl_sync(0xffffffffu, fc, 0);
        // pass 2: emit the groups in key order
        uint32_t mask = 0, ord0 = 0, gs_carry = l2;
        for (uint32_t b0 = l2; b0 < hi; b0 += 32) {
          const uint32_t j = b0 + lane;
          const bool bd = j < hi && (j + 1 == hi || lcp[j] == d);
          const uint32_t bal = __ballot_sync(0xffffffffu, bd);
          if (bd) {
            const uint32_t below = bal & ((1u << lane) - 1u);
            const uint32_t gs = below ? b0 + (31u - __clz(below)) + 1u : gs_carry;
            mask |= 1u << key_nib(v, gs, d, in.key_bytes, koff);
            R[fc + ord0 + __popc(below)] = make_uint4(gs | ((j + 1) << 16), d + 1, 0, kPending);
          }
          if (bal) gs_carry = b0 + (31u - __clz(bal)) + 1u;
          ord0 += __popc(bal);
        }
        for (int o = 16; o; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
        r.x = kTBranch | (d << 16);
        r.z = base + fc;
        r.w = mask;
      }
      if (lane == 0) R[e] = r;
    }
    __syncthreads();
    head = tail;
    rounds++;
  }
  const uint32_t count = s_tail;

  // ---- leaf lengths (parallel: one global value_len read per leaf, all in flight at once)
  for (uint32_t id = tid; id < count; id += kTrieThreads) {
    const uint4 r = R[id];
    if (rec_kind(r) == kTLeaf) {
      const uint32_t vl = in.value_len[r.y];
      const uint32_t v0 = vl == 1 ? in.value_bytes[in.value_off[r.y]] : 0u;
      const uint32_t payload = hp_item_size(rec_pl(r)) + str_item_size(vl, v0);
      Ln[id] = hdr_size(payload) + payload;
    }
  }
  __syncthreads();

  // ---- inner nodes, rounds in reverse (children before parents), one thread per node: length, height
  for (uint32_t rd = rounds; rd-- > 0;) {
    const uint32_t r0 = s_round[rd], r1 = s_round[rd + 1];
    for (uint32_t id = r0 + tid; id < r1; id += kTrieThreads) {
      uint4 r = R[id];
      if (rec_kind(r) == kTLeaf) continue;
      uint32_t payload, h = 0;
      const uint32_t c0 = r.z - base;
      if (rec_kind(r) == kTExt) {
        payload = hp_item_size(rec_pl(r)) + ref_size(Ln[c0]);
        h = rec_height(R[c0]) + 1;
      } else {
        const uint32_t nc = __popc(rec_mask(r));
        payload = 16 - nc;
        for (uint32_t c = 0; c < nc; c++) {
          payload += ref_size(Ln[c0 + c]);
          h = max(h, rec_height(R[c0 + c]) + 1);
        }
        if (r.y == kNoItem) payload += 1;
        else {
          const uint32_t vl = in.value_len[r.y];
          payload += str_item_size(vl, vl == 1 ? in.value_bytes[in.value_off[r.y]] : 0u);
        }
      }
      r.x |= h << 8;
      R[id] = r;
      Ln[id] = hdr_size(payload) + payload;
    }
    __syncthreads();
  }

  // ---- hashed flag, level histogram, Keccak-f count, arena offsets (block scan over id chunks)
  {
    const uint32_t seg = (count + kTrieThreads - 1) / kTrieThreads;
    const uint32_t s0 = min(count, tid * seg), s1 = min(count, s0 + seg);
    uint32_t bytes = 0, perms = 0, nh = 0;
    for (uint32_t id = s0; id < s1; id++) {
      const uint32_t len = Ln[id];
      const uint32_t hashed = (len >= 32 || id == 0) ? 1u : 0u;  // write_node: >= 32 bytes by hash; the root always
      uint4 r = R[id];
      r.w |= hashed << 16;
      R[id] = r;
      bytes += (len + 15u) & ~15u;
      if (hashed) { perms += len / 136u + 1u; nh++; }
      atomicAdd(&s_hist[2 * rec_height(r) + (hashed ? 0 : 1)], 1u);
    }
    s_scan[tid] = bytes;
    if (perms) { atomicAdd(&s_perms, (unsigned long long)perms); atomicAdd(&s_hashed, (unsigned long long)nh); }
    __syncthreads();
    if (tid == 0) {
      unsigned long long acc = 0;
      for (uint32_t i = 0; i < kTrieThreads; i++) { const uint32_t c = s_scan[i]; s_scan[i] = (uint32_t)acc; acc += c; }
      w.tcount[t] = count;
      s_arena_base = atomicAdd(&w.sum->arena_bytes, acc);
      atomicAdd(&w.sum->perms, s_perms);
      atomicAdd(&w.sum->nodes_hashed, s_hashed);
      atomicAdd(&w.sum->n_nodes, (unsigned long long)count);
    }
    __syncthreads();
    unsigned long long o = s_arena_base + s_scan[tid];
    for (uint32_t id = s0; id < s1; id++) {
      w.off[base + id] = o;
      o += (Ln[id] + 15u) & ~15u;
    }
  }
  if (in_smem) {
    for (uint32_t id = tid; id < count; id += kTrieThreads) { w.rec[base + id] = R[id]; w.len[base + id] = Ln[id]; }
  }
  for (uint32_t i = tid; i < 2 * kMaxLevels; i += kTrieThreads)
    if (s_hist[i]) atomicAdd(&w.sum->lvl_count[i], s_hist[i]);
}

// ------------------------------------------------------------------ level lists
__global__ void k_trie_level_scan(TrieSummary* sum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t acc = 0;
    for (int i = 0; i < 2 * kMaxLevels; i++) { sum->lvl_cursor[i] = acc; acc += sum->lvl_count[i]; }
  }
}

__global__ void __launch_bounds__(128) k_trie_level_scatter(const TrieBatchDev in, const TrieWork w) {
  __shared__ uint32_t cnt[2 * kMaxLevels];
  __shared__ uint32_t start[2 * kMaxLevels];
  const uint32_t t = blockIdx.x, tid = threadIdx.x;
  const uint32_t count = w.tcount[t];
  if (count == 0) return;
  const uint32_t base = 3u * in.trie_first[t];
  for (uint32_t i = tid; i < 2 * kMaxLevels; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (uint32_t id = tid; id < count; id += blockDim.x) {
    const uint4 r = w.rec[base + id];
    atomicAdd(&cnt[2 * rec_height(r) + (rec_hashed(r) ? 0 : 1)], 1u);
  }
  __syncthreads();
  for (uint32_t i = tid; i < 2 * kMaxLevels; i += blockDim.x) {
    if (cnt[i]) start[i] = atomicAdd(&w.sum->lvl_cursor[i], cnt[i]);
    cnt[i] = 0;
  }
  __syncthreads();
  for (uint32_t id = tid; id < count; id += blockDim.x) {
    const uint4 r = w.rec[base + id];
    const uint32_t k = 2 * rec_height(r) + (rec_hashed(r) ? 0 : 1);
    w.lvl_list[start[k] + atomicAdd(&cnt[k], 1u)] = base + id;
  }
}

// ------------------------------------------------------------------ encode
namespace {

// warp-cooperative copy of n bytes from a 4-byte aligned src to an arbitrarily aligned dst;
// src must be readable up to the next multiple of 4 after src + n
__device__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
  uint32_t hb = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
  if (hb > n) hb = n;
  if (lane < hb) dst[lane] = src[lane];
  const uint32_t nw = (n - hb) >> 2;
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + hb);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
  const uint32_t sh = 8u * hb;  // src byte hb + 4w sits hb bytes into word w
  for (uint32_t k = lane; k < nw; k += 32) {
    const uint32_t a = __ldg(sw + k);
    const uint32_t b = hb ? __ldg(sw + k + 1) : 0u;
    dw[k] = __funnelshift_r(a, b, sh);
  }
  const uint32_t done = hb + 4u * nw;
  if (lane < n - done) dst[done + lane] = src[done + lane];
}

// RLP string item for a value (leaf item 1 / branch item 16), written by the whole warp
__device__ uint32_t warp_put_value(uint8_t* dst, const uint8_t* val, uint32_t vl, uint32_t lane) {
  const uint32_t v0 = vl ? val[0] : 0u;
  if (vl == 1 && v0 < 0x80) {
    if (lane == 0) dst[0] = (uint8_t)v0;
    return 1;
  }
  const uint32_t h = hdr_size(vl);
  if (lane == 0) put_hdr(dst, vl, false);
  warp_copy(dst + h, val, vl, lane);
  return h + vl;
}

}  // namespace

// one warp writes the RLP encoding of `node` to dst (children are read from `arena` / w.digests)
constexpr uint32_t kEncStage = 576;  // per-warp shared staging: a branch without a value is at most 3 + 16 * 33 + 1 bytes

// sbuf: kEncStage bytes of shared memory owned by this warp (16-byte aligned)
__device__ void warp_encode_node(const TrieBatchDev& in, const TrieWork& w, const uint8_t* __restrict__ arena, uint32_t node,
                                 uint8_t* __restrict__ dst, uint32_t lane, uint8_t* sbuf) {
  const uint4 r = w.rec[node];
  const uint32_t len = w.len[node];
  const uint32_t kind = rec_kind(r);
  if (kind == kTLeaf || kind == kTExt) {
    const uint8_t* key = in.key_bytes + in.key_off[r.y];
    const uint32_t ps = rec_ps(r), pl = rec_pl(r);
    const uint32_t hpn = pl / 2 + 1;
    const uint32_t pis = hp_item_size(pl);
    const uint32_t payload = payload_of(len);
    const uint32_t hl = len - payload;
    if (lane == 0) {
      put_hdr(dst, payload, true);
      if (hpn > 1) dst[hl] = (uint8_t)(0x80 + hpn);
    }
    uint8_t* pp = dst + hl + (hpn > 1 ? 1 : 0);
    for (uint32_t i = lane; i < hpn; i += 32) pp[i] = (uint8_t)hp_byte(key, ps, pl, kind == kTLeaf, i);
    uint8_t* vp = dst + hl + pis;
    if (kind == kTLeaf) {
      warp_put_value(vp, in.value_bytes + in.value_off[r.y], in.value_len[r.y], lane);
    } else {
      const uint32_t cl = w.len[r.z];
      if (cl < 32) { if (lane < cl) vp[lane] = arena[w.off[r.z] + lane]; }
      else {
        if (lane == 0) vp[0] = 0xa0;
        vp[1 + lane] = w.digests[32ull * r.z + lane];
      }
    }
  } else {
    const uint32_t mask = rec_mask(r);
    const uint32_t payload = payload_of(len);
    const uint32_t hl = len - payload;
    // lanes 0..15: child slots; lane 16: value item
    uint32_t sz = 0, child = 0, cl = 0;
    const bool has = lane < 16 && ((mask >> lane) & 1u);
    if (lane < 16) {
      if (has) { child = r.z + __popc(mask & ((1u << lane) - 1u)); cl = w.len[child]; sz = ref_size(cl); }
      else sz = 1;
    }
    uint32_t vl = 0;
    const uint8_t* val = nullptr;
    if (r.y != kNoItem) { vl = in.value_len[r.y]; val = in.value_bytes + in.value_off[r.y]; }
    if (lane == 16) sz = (r.y == kNoItem) ? 1u : str_item_size(vl, vl ? val[0] : 0u);
    uint32_t incl = sz;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += x;
    }
    // A branch without a value (every inner node of a tx / receipt / state trie) is assembled in shared
    // memory and leaves as coalesced 16-byte stores; writing the 33-byte items straight to global memory
    // costs one scattered byte store per lane per byte (30 x the sector writes).
    const bool staged = r.y == kNoItem && len <= kEncStage;
    uint8_t* out = staged ? sbuf : dst;
    uint8_t* ip = out + hl + (incl - sz);
    if (lane == 0) put_hdr(out, payload, true);
    if (lane < 16) {
      if (!has) ip[0] = 0x80;
      else if (cl < 32) { const uint8_t* s = arena + w.off[child]; for (uint32_t i = 0; i < cl; i++) ip[i] = s[i]; }
      else {
        ip[0] = 0xa0;
        const uint4* s = reinterpret_cast<const uint4*>(w.digests + 32ull * child);
        const uint4 a = s[0], b = s[1];
        const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 32; i++) ip[1 + i] = (uint8_t)(ws[i >> 2] >> (8 * (i & 3)));
      }
    }
    const uint32_t voff = __shfl_sync(0xffffffffu, incl - sz, 16);
    if (r.y == kNoItem) { if (lane == 16) ip[0] = 0x80; }
    else warp_put_value(dst + hl + voff, val, vl, lane);
    if (staged) {
      __syncwarp();
      const uint4* s4 = reinterpret_cast<const uint4*>(sbuf);
      uint4* d4 = reinterpret_cast<uint4*>(dst);  // node slots are 16-byte aligned and padded to 16
      for (uint32_t c = lane; c < (len + 15u) / 16u; c += 32) d4[c] = s4[c];
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(256) k_trie_encode(const TrieBatchDev in, const TrieWork w, const uint32_t* __restrict__ list,
                                                     uint32_t n_list, uint8_t* __restrict__ arena) {
  const uint32_t lane = threadIdx.x & 31u;
  __shared__ __align__(16) uint8_t s_enc[8][kEncStage];
  const uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wi >= n_list) return;
  const uint32_t node = list[wi];
  warp_encode_node(in, w, arena, node, arena + w.off[node], lane, s_enc[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(256) k_trie_roots(const TrieBatchDev in, const TrieWork w, uint8_t* __restrict__ roots32) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t t = i >> 3;
  const uint32_t part = (uint32_t)i & 7u;
  if (t >= in.n_tries) return;
  // keccak256(rlp("")) -- EthTrie::new's root before any insert
  const uint32_t kEmpty[8] = {0x171fe856u, 0xa655cc1bu, 0xe64583ffu, 0x6ef8c092u, 0x1be0485bu, 0xc0ad6c99u, 0xb52f6201u, 0x21b463e3u};
  uint32_t x = kEmpty[part];
  if (w.tcount[t]) x = reinterpret_cast<const uint32_t*>(w.digests + 32ull * (3ull * in.trie_first[t]))[part];
  reinterpret_cast<uint32_t*>(roots32 + 32 * t)[part] = x;
}

// ------------------------------------------------------------------ host launchers
size_t trie_structure_smem(uint32_t cap, uint32_t node_cap) { return ((14ull * cap + 15) & ~15ull) + 20ull * node_cap; }
// node records live in shared memory while the CTA's footprint stays small enough for several CTAs per SM
constexpr uint32_t kTrieSmemNodeItems = 1024;

cudaError_t trie_init_device() {
  return cudaFuncSetAttribute(k_trie_structure, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)trie_structure_smem(kTrieMaxItems, 3 * kTrieSmemNodeItems));
}

cudaError_t launch_trie_scan_input(const TrieBatchDev& in, TrieSummary* sum, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(TrieSummary), st);
  if (e != cudaSuccess) return e;
  const uint64_t n = in.n_items > in.n_tries ? in.n_items : in.n_tries;
  if (n == 0) return cudaSuccess;
  k_trie_scan_input<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, sum);
  return cudaGetLastError();
}

cudaError_t launch_trie_structure(const TrieBatchDev& in, const TrieWork& w, uint32_t max_items, cudaStream_t st) {
  if (in.n_tries == 0) return cudaSuccess;
  uint32_t cap = 2;
  while (cap < max_items) cap <<= 1;
  const uint32_t node_cap = 3 * std::min(max_items, kTrieSmemNodeItems);  // larger tries use the global table
  k_trie_structure<<<in.n_tries, kTrieThreads, trie_structure_smem(cap, node_cap), st>>>(in, w, cap, node_cap);
  k_trie_level_scan<<<1, 32, 0, st>>>(w.sum);
  k_trie_level_scatter<<<in.n_tries, 128, 0, st>>>(in, w);
  return cudaGetLastError();
}

cudaError_t launch_trie_encode(const TrieBatchDev& in, const TrieWork& w, const uint32_t* list, uint32_t n_list,
                               uint8_t* arena, cudaStream_t st) {
  if (n_list == 0) return cudaSuccess;
  const unsigned blocks = (n_list + 7) / 8;
  k_trie_encode<<<blocks, 256, 0, st>>>(in, w, list, n_list, arena);
  return cudaGetLastError();
}

cudaError_t launch_trie_roots(const TrieBatchDev& in, const TrieWork& w, uint8_t* roots32, cudaStream_t st) {
  if (in.n_tries == 0) return cudaSuccess;
  const uint64_t threads = (uint64_t)in.n_tries * 8;
  k_trie_roots<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, w, roots32);
  return cudaGetLastError();
}

}  // namespace mptv

// ------------------------------------------------------------------ get_proof (after a rebuild)
// Trie::get_proof(key) of eth_trie for targets (trie, key): the encoded nodes on the key's path,
// root first -- the root always, every other node only when its parent references it by hash
// (>= 32 bytes); inline nodes are part of their parent.  The walk stops at a leaf, at an
// extension whose path the key does not share, at an empty branch slot or when the key is
// exhausted (proof of absence).  /root/reference/trie-utils/src/proofs/transaction.rs:66-73,
// proofs/receipt.rs:84-92 (trie.get_proof(&key)).
namespace mptv {

namespace {

// visits the path; fn(node id) for every emitted node
template <class F>
__device__ void proof_walk(const TrieBatchDev& in, const TrieWork& w, uint32_t trie, const uint8_t* key, uint32_t klen, F fn) {
  if (w.tcount[trie] == 0) return;
  uint32_t node = 3u * in.trie_first[trie];
  uint32_t idx = 0;
  const uint32_t plen = 2 * klen;
  for (;;) {
    const uint4 r = w.rec[node];
    if ((r.w >> 16) & 1u) fn(node);
    const uint32_t kind = r.x & 0xffu;
    if (kind == kTLeaf) return;
    if (kind == kTExt) {
      const uint32_t ps = (r.x >> 16) & 0xffu, pl = r.x >> 24;
      if (plen - idx < pl) return;
      const uint8_t* ek = in.key_bytes + in.key_off[r.y];
      for (uint32_t i = 0; i < pl; i++) {
        const uint32_t a = ek[(ps + i) >> 1], b = key[(idx + i) >> 1];
        const uint32_t na = ((ps + i) & 1) ? (a & 15u) : (a >> 4), nb = ((idx + i) & 1) ? (b & 15u) : (b >> 4);
        if (na != nb) return;
      }
      idx += pl;
      node = r.z;
    } else {
      if (idx >= plen) return;
      const uint32_t b = key[idx >> 1];
      const uint32_t nib = (idx & 1) ? (b & 15u) : (b >> 4);
      const uint32_t mask = r.w & 0xffffu;
      if (!((mask >> nib) & 1u)) return;
      node = r.z + __popc(mask & ((1u << nib) - 1u));
      idx++;
    }
  }
}
}  // namespace

// pass 1: nodes and (16-byte padded) bytes per target
__global__ void __launch_bounds__(256) k_trie_proof_count(const TrieBatchDev in, const TrieWork w, const uint32_t* __restrict__ target_trie,
                                                          const uint8_t* __restrict__ tkey_bytes, const uint32_t* __restrict__ tkey_off,
                                                          uint32_t n_targets, uint32_t* __restrict__ cnt, uint64_t* __restrict__ bytes) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_targets) return;
  uint32_t c = 0;
  uint64_t b = 0;
  proof_walk(in, w, target_trie[q], tkey_bytes + tkey_off[q], tkey_off[q + 1] - tkey_off[q],
             [&](uint32_t node) { c++; b += (w.len[node] + 15u) & ~15u; });
  cnt[q] = c;
  bytes[q] = b;
}

// exclusive scan of cnt -> proof_first[n+1] and bytes -> byte_first[n+1]; one CTA, chunk per thread
__global__ void __launch_bounds__(1024) k_trie_proof_scan(const uint32_t* __restrict__ cnt, const uint64_t* __restrict__ bytes,
                                                          uint32_t n, uint32_t* __restrict__ proof_first,
                                                          uint64_t* __restrict__ byte_first) {
  __shared__ unsigned long long sc[1024], sb[1024];
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t s0 = min(n, tid * per), s1 = min(n, s0 + per);
  unsigned long long c = 0, b = 0;
  for (uint32_t i = s0; i < s1; i++) { c += cnt[i]; b += bytes[i]; }
  sc[tid] = c; sb[tid] = b;
  __syncthreads();
  if (tid == 0) {
    unsigned long long ac = 0, ab = 0;
    for (int i = 0; i < 1024; i++) {
      const unsigned long long x = sc[i], y = sb[i];
      sc[i] = ac; sb[i] = ab;
      ac += x; ab += y;
    }
    proof_first[n] = (uint32_t)ac;
    byte_first[n] = ab;
  }
  __syncthreads();
  c = sc[tid]; b = sb[tid];
  for (uint32_t i = s0; i < s1; i++) {
    proof_first[i] = (uint32_t)c; byte_first[i] = b;
    c += cnt[i]; b += bytes[i];
  }
}

// pass 2: one warp per target copies its nodes into the output arena
__global__ void __launch_bounds__(256) k_trie_proof_emit(const TrieBatchDev in, const TrieWork w, const uint8_t* __restrict__ arena,
                                                         const uint32_t* __restrict__ target_trie, const uint8_t* __restrict__ tkey_bytes,
                                                         const uint32_t* __restrict__ tkey_off, uint32_t n_targets,
                                                         const uint32_t* __restrict__ proof_first, const uint64_t* __restrict__ byte_first,
                                                         uint8_t* __restrict__ out_bytes, uint64_t* __restrict__ out_off,
                                                         uint32_t* __restrict__ out_len, int leaves_in_arena) {
  __shared__ __align__(16) uint8_t s_enc[8][kEncStage];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_targets) return;
  uint32_t k = proof_first[q];
  uint64_t o = byte_first[q];
  proof_walk(in, w, target_trie[q], tkey_bytes + tkey_off[q], tkey_off[q + 1] - tkey_off[q], [&](uint32_t node) {
    const uint32_t len = w.len[node];
    if (!leaves_in_arena && (w.rec[node].x & 0xffu) == kTLeaf) {
      // hashed leaves were digested straight from the value arena (K1L): encode this one now
      warp_encode_node(in, w, arena, node, out_bytes + o, lane, s_enc[threadIdx.x >> 5]);
    } else {
      const uint4* s = reinterpret_cast<const uint4*>(arena + w.off[node]);  // 16-byte aligned slots on both sides
      uint4* d = reinterpret_cast<uint4*>(out_bytes + o);
      for (uint32_t i = lane; i < (len + 15u) / 16u; i += 32) {
        uint4 x = s[i];
        const int r = (int)len - 16 * (int)i;  // the slot's padding in the arena is whatever an earlier call left: emit zeros
        if (r < 16) {
          uint32_t wd[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int keep = r - 4 * c;
            wd[c] &= keep >= 4 ? 0xffffffffu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
          }
          x = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
        d[i] = x;
      }
    }
    if (lane == 0) { out_off[k] = o; out_len[k] = len; }
    k++;
    o += (len + 15u) & ~15u;
  });
}

cudaError_t launch_trie_proof_count(const TrieBatchDev& in, const TrieWork& w, const uint32_t* target_trie,
                                    const uint8_t* tkey_bytes, const uint32_t* tkey_off, uint32_t n_targets, uint32_t* cnt,
                                    uint64_t* bytes, uint32_t* proof_first, uint64_t* byte_first, cudaStream_t st) {
  if (n_targets == 0) return cudaSuccess;
  k_trie_proof_count<<<(n_targets + 255) / 256, 256, 0, st>>>(in, w, target_trie, tkey_bytes, tkey_off, n_targets, cnt, bytes);
  k_trie_proof_scan<<<1, 1024, 0, st>>>(cnt, bytes, n_targets, proof_first, byte_first);
  return cudaGetLastError();
}

cudaError_t launch_trie_proof_emit(const TrieBatchDev& in, const TrieWork& w, const uint8_t* arena, const uint32_t* target_trie,
                                   const uint8_t* tkey_bytes, const uint32_t* tkey_off, uint32_t n_targets,
                                   const uint32_t* proof_first, const uint64_t* byte_first, uint8_t* out_bytes,
                                   uint64_t* out_off, uint32_t* out_len, bool leaves_in_arena, cudaStream_t st) {
  if (n_targets == 0) return cudaSuccess;
  k_trie_proof_emit<<<(n_targets + 7) / 8, 256, 0, st>>>(in, w, arena, target_trie, tkey_bytes, tkey_off, n_targets, proof_first,
                                                         byte_first, out_bytes, out_off, out_len, leaves_in_arena ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace mptv
