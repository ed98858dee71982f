// verify_device.cuh -- internal: the device-side rule set of the verifier, shared by the batch kernels
// (verify_kernels.cu: K2a / K2f / K2b over device-resident arenas) and the single-launch latency kernel
// (single_kernels.cu: the same functions over a batch staged in shared memory).
//
// What these functions restate is crypto_ops::verify_merkle_proof after hashing
// (/root/reference/crypto-ops/src/lib.rs:14-22) and the third-party code under it (eth_trie@ade617b decode_node /
// get_at / write_node, alloy-rlp Header::decode): rules R1..R20 of SURVEY.md Appendix A.
//
// MPTV_LDG(p) is how read-only input is loaded.  The batch kernels define it as __ldg (ld.global.nc: the arenas are
// never written while they run); the latency kernel reads a copy it made itself earlier in the same launch, in
// shared memory, and uses plain loads.  Everything here has internal linkage: each including .cu gets its own copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

#ifndef MPTV_LDG
#define MPTV_LDG(p) __ldg(p)
#endif
#ifndef MPTV_WALK_MINB
#define MPTV_WALK_MINB 4  // measured on B200: 4 x 256 threads / SM beats 2, 3, 5, 6, 8 (tools/ sweep, r01)
#endif

namespace mptv {
namespace {

// ------------------------------------------------------------------ RLP header (alloy-rlp, R17)
struct Hdr { uint32_t is_list, hdr_len, payload_len; };

__device__ __forceinline__ uint32_t ldb(const uint8_t* p) { return (uint32_t)MPTV_LDG(p); }

// strict canonical header of the item that starts at p with n bytes available
__device__ bool rlp_hdr(const uint8_t* p, uint32_t n, Hdr& h) {
  if (n == 0) return false;
  uint32_t b = ldb(p);
  if (b < 0x80) { h.is_list = 0; h.hdr_len = 0; h.payload_len = 1; return true; }
  if (b < 0xB8) {
    h.is_list = 0; h.hdr_len = 1; h.payload_len = b - 0x80;
    if (h.payload_len == 1) {
      if (n < 2) return false;
      if (ldb(p + 1) < 0x80) return false;  // NonCanonicalSingleByte
    }
  } else if (b < 0xC0 || b >= 0xF8) {
    h.is_list = b >= 0xF8 ? 1u : 0u;
    uint32_t ll = h.is_list ? b - 0xF7 : b - 0xB7;
    if (n < 1 + ll) return false;
    if (ldb(p + 1) == 0) return false;  // LeadingZero
    if (ll > 4) return false;           // cannot fit in a u32-sized node
    uint32_t v = 0;
    for (uint32_t i = 0; i < ll; i++) v = (v << 8) | ldb(p + 1 + i);
    if (v < 56) return false;  // NonCanonicalSize
    h.hdr_len = 1 + ll; h.payload_len = v;
  } else {
    h.is_list = 1; h.hdr_len = 1; h.payload_len = b - 0xC0;
  }
  return (uint64_t)h.hdr_len + h.payload_len <= (uint64_t)n;  // InputTooShort
}

// ------------------------------------------------------------------ K2a: per-node decode
// Frame of the explicit DFS stack over nested inline nodes.
struct Frame { uint32_t pos, end; uint8_t cnt, idx, leaf, top; };

// scan the items of the list at p[lst .. lst+hdr+payload): every header must be valid and fit;
// returns the item count (18 means "more than 17") or -1 on a header error
__device__ int scan_items(const uint8_t* p, uint32_t lst, const Hdr& h) {
  uint32_t q = lst + h.hdr_len, e = q + h.payload_len;
  int cnt = 0;
  while (q < e) {
    Hdr t;
    if (!rlp_hdr(p + q, e - q, t)) return -1;
    q += t.hdr_len + t.payload_len;
    if (++cnt > 17) return 18;
  }
  return cnt;
}

__device__ __forceinline__ bool value_item_canonical(const Hdr& t) {
  // R20 x R4: the value is re-encoded as an RLP string of the decoded value bytes; that reproduces
  // the original item iff it was a string and not the 2-byte form 0x81 b (decoded as [0x81, b]).
  return !t.is_list && !(t.hdr_len == 1 && t.payload_len == 1);
}

// The DFS keeps the most recent kInlineWindow frames (a circular window indexed by absolute depth).  Inline
// extensions chain as tail calls and use no frame; only inline nodes under a BRANCH do.  When the walk returns to a
// frame that has fallen out of the window -- inline branches nested deeper than 64 levels: >= 1.1 KB of
// purpose-built bytes, an inline node being < 32 bytes in any real trie -- the window is rebuilt by walking down
// again from the top of the node along the path to the child that just ended (`cend` = its end).  Every list on
// that path was validated on the way down, so the replay only skips over sibling headers: 17 header decodes a level,
// once per 64 levels of unwinding.  No depth limit, no scratch memory: the reference recurses without a limit too
// (and is itself quadratic in the depth).
__device__ int replay_frames(const uint8_t* p, const Hdr& top, uint32_t cend, Frame* st, int& base) {
  uint32_t lst = 0;
  Hdr lh = top;
  int d = 0;
  for (;;) {
    const int cnt = scan_items(p, lst, lh);  // 2 or 17: validated when the walk first came through here
    const uint32_t lend = lst + lh.hdr_len + lh.payload_len;
    uint32_t q = lst + lh.hdr_len, e = q, idx = 0;
    Hdr t;
    for (;; idx++) {  // the item that contains byte cend - 1
      rlp_hdr(p + q, lend - q, t);
      e = q + t.hdr_len + t.payload_len;
      if (cend <= e) break;
      q = e;
    }
    if (cnt == 17) {
      // a branch on the path: its frame as it was when the walk descended into this child
      Frame& f = st[d & (kInlineWindow - 1)];
      f.pos = e; f.end = lend; f.cnt = 17; f.idx = (uint8_t)(idx + 1); f.leaf = 0; f.top = lst == 0 ? 1 : 0;
      if (e == cend) break;  // ... and this one is the parent of the child that ended
      d++;
    }
    lst = q;  // descend (a 2-item list is an extension: its child shares the frame, the depth stays)
    lh = t;
  }
  base = d >= kInlineWindow ? d - kInlineWindow + 1 : 0;
  return d;
}

__device__ uint32_t parse_node(const uint8_t* p, uint32_t n) {
  Hdr h;
  if (!rlp_hdr(p, n, h)) return make_meta(kKindEmpty, kDecErr, 0, 0, 0, 0);
  uint32_t canon = (h.hdr_len + h.payload_len == n) ? 1u : 0u;  // trailing bytes (R18 vs R4)
  if (!h.is_list) {
    if (h.payload_len == 0) return make_meta(kKindEmpty, kDecOk, canon, 0, h.hdr_len, 0);
    if (h.payload_len == 32) {
      // the reference converts EVERYTHING after the header to a B256 (FixedBytes::from_slice): a bare
      // 32-byte string followed by anything is a raw panic, not a hash node with ignored trailing bytes
      if (n != h.hdr_len + 32u) return make_meta(kKindEmpty, kDecPanic, 0, 0, 0, 0);
      return make_meta(kKindHash, kDecOk, 0, 0, h.hdr_len, 0);
    }
    return make_meta(kKindEmpty, kDecErr, 0, 0, 0, 0);
  }
  Frame st[kInlineWindow];
  int depth = 0, base = 0;  // absolute depth of the current frame / of the oldest frame still in the window
  {
    int c = scan_items(p, 0, h);
    if (c != 2 && c != 17) return make_meta(kKindEmpty, kDecErr, 0, 0, 0, 0);
    st[0].pos = h.hdr_len; st[0].end = h.hdr_len + h.payload_len;
    st[0].cnt = (uint8_t)c; st[0].idx = 0; st[0].leaf = 0; st[0].top = 1;
  }
  uint32_t top_kind = st[0].cnt == 17 ? kKindBranch : kKindExt;
  uint32_t mask = 0, fast = st[0].cnt == 17 ? 1u : 0u;
  for (;;) {
    Frame& f = st[depth & (kInlineWindow - 1)];
    if (f.idx == f.cnt) {
      if (depth == 0) break;
      const uint32_t cend = f.end;  // the list that just ended
      depth--;
      if (depth < base) depth = replay_frames(p, h, cend, st, base);
      continue;
    }
    Hdr t;
    rlp_hdr(p + f.pos, f.end - f.pos, t);  // validated by scan_items
    const uint32_t item = f.pos;
    const uint32_t i = f.idx;
    f.pos += t.hdr_len + t.payload_len;
    f.idx++;
    bool is_child = false;
    if (f.cnt == 2) {
      if (i == 0) {
        // Nibbles::from_compact on the item's payload (R19); list/string flag not checked
        if (t.payload_len == 0) return make_meta(kKindEmpty, kDecPanic, 0, 0, 0, 0);
        uint32_t b = ldb(p + item + t.hdr_len);
        uint32_t flag = b >> 4;
        if (flag > 3) return make_meta(kKindEmpty, kDecPanic, 0, 0, 0, 0);
        uint32_t nn = (t.payload_len - 1) * 2 + (flag & 1);
        f.leaf = flag >= 2;
        // decode_node calls key.is_leaf() = hex_data[len-1]: panics on an empty extension path
        if (!f.leaf && nn == 0) return make_meta(kKindEmpty, kDecPanic, 0, 0, 0, 0);
        if (t.is_list || (!(flag & 1) && (b & 15))) canon = 0;
        if (f.top) top_kind = f.leaf ? kKindLeaf : kKindExt;
      } else if (f.leaf) {
        if (!value_item_canonical(t)) canon = 0;
      } else {
        is_child = true;
      }
    } else {
      if (i < 16) is_child = true;
      else {
        if (!value_item_canonical(t)) canon = 0;
        // "fast" (plain branch) promises the walk an EMPTY value item, i.e. exactly 0x80
        if (f.top && (t.is_list || t.payload_len != 0)) fast = 0;
      }
    }
    if (is_child) {
      if (t.is_list) {
        // inline node: decoded recursively; re-encodes in place only while < 32 bytes (write_node)
        if (t.hdr_len + t.payload_len >= 32) canon = 0;
        if (f.top) fast = 0;
        int c = scan_items(p, item, t);
        if (c != 2 && c != 17) return make_meta(kKindEmpty, kDecErr, 0, 0, 0, 0);
        // An extension's child is the LAST item of its list: nothing of the parent is left to visit, so
        // the child takes over the parent's frame (a tail call).  Only inline nodes under a BRANCH take a frame.
        if (f.cnt != 2) {
          depth++;
          if (depth - base >= kInlineWindow) base++;  // the oldest frame falls out of the window (rebuilt on return)
        }
        Frame& g = st[depth & (kInlineWindow - 1)];
        g.pos = item + t.hdr_len; g.end = item + t.hdr_len + t.payload_len;
        g.cnt = (uint8_t)c; g.idx = 0; g.leaf = 0; g.top = 0;
      } else if (t.payload_len == 32) {
        if (f.top) mask |= 1u << i;
      } else if (t.payload_len != 0) {
        return make_meta(kKindEmpty, kDecErr, 0, 0, 0, 0);  // InvalidData (R16)
      }
    }
  }
  return make_meta(top_kind, kDecOk, canon, fast, h.hdr_len, mask);
}

// ------------------------------------------------------------------ K2b: the walk
// alloy_rlp::decode_exact::<Account> (storage-circuit/src/main.rs:15): rlp([nonce u64, balance
// U256, storage_root B256, code_hash B256]) with nothing left over.  Returns the offset of the
// 32-byte storage_root inside v, or 0xffffffff.
__device__ uint32_t account_storage_root_off(const uint8_t* v, uint32_t n) {
  Hdr h, t;
  if (!rlp_hdr(v, n, h) || !h.is_list) return 0xffffffffu;
  if (h.hdr_len + h.payload_len != n) return 0xffffffffu;
  uint32_t q = h.hdr_len, e = n, off = 0xffffffffu;
  for (int i = 0; i < 4; i++) {
    if (!rlp_hdr(v + q, e - q, t) || t.is_list) return 0xffffffffu;
    if (i == 0 && t.payload_len > 8) return 0xffffffffu;
    if (i == 1 && t.payload_len > 32) return 0xffffffffu;
    if (i < 2 && t.payload_len > 0 && ldb(v + q + t.hdr_len) == 0) return 0xffffffffu;
    if (i >= 2 && t.payload_len != 32) return 0xffffffffu;
    if (i == 2) off = q + t.hdr_len;
    q += t.hdr_len + t.payload_len;
  }
  return q == e ? off : 0xffffffffu;
}

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
  return ldb(p) | (ldb(p + 1) << 8) | (ldb(p + 2) << 16) | (ldb(p + 3) << 24);
}

template <int G>
struct Group {
  uint32_t gmask;   // lanes of this group within the warp
  uint32_t gshift;  // first lane of the group
  uint32_t lig;     // lane in group
  __device__ Group() {
    uint32_t lane = threadIdx.x & 31;
    gshift = lane & ~(uint32_t)(G - 1);
    lig = lane & (G - 1);
    gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gshift;
  }
  __device__ __forceinline__ uint32_t ballot(bool p) const {
    return (__ballot_sync(gmask, p) & gmask) >> gshift;
  }
  __device__ __forceinline__ bool all(bool p) const { return __all_sync(gmask, p); }
  __device__ __forceinline__ uint32_t bcast(uint32_t v, uint32_t src) const {
    return __shfl_sync(gmask, v, src, G);
  }
};

// key nibble at path index i; the terminator 16 sits at i == 2*klen (Nibbles::from_raw, R11)
__device__ __forceinline__ uint32_t key_nibble(const uint8_t* key, uint32_t klen, uint32_t i) {
  if (i >= 2 * klen) return 16;
  uint32_t b = ldb(key + (i >> 1));
  return (i & 1) ? (b & 15) : (b >> 4);
}

template <int G>
__device__ void walk_one(const DeviceBatch& b, int wave, const Group<G>& g, const uint64_t p,
                         const uint8_t* __restrict__ digests, const uint32_t* __restrict__ meta, uint8_t* status_out,
                         uint64_t* value_off_out, uint32_t* value_len_out) {
  const uint8_t* __restrict__ node_bytes = b.node_bytes - b.byte_base;  // indexed by GLOBAL offsets
  const uint64_t* __restrict__ node_off = b.node_off;
  const uint32_t* __restrict__ node_len = b.node_len;
  const uint32_t* __restrict__ proof_first = b.proof_first;
  const int32_t* __restrict__ root_from_proof = b.root_from_proof;
  const bool dependent = root_from_proof != nullptr && root_from_proof[p] >= 0;
  if (dependent != (wave == 1)) return;

  const uint32_t a = proof_first[p] - b.node_base, n = proof_first[p + 1] - proof_first[p];
  const uint8_t* key = b.key_bytes + (b.key_off[p] - b.key_base);
  const uint32_t klen = b.key_len ? b.key_len[p] : b.key_off[p + 1] - b.key_off[p];

  uint32_t status = kStOk;
  uint64_t voff = 0;
  uint32_t vlen = 0;

  // ---- where the 32-byte root lives
  const uint8_t* rp = b.roots + 32 * p;
  if (dependent) {
    // nested workload (storage-circuit/src/main.rs:10-27): root = storage_root of the account
    // leaf proven by proof d, which wave 0 has already judged
    const uint64_t d = (uint64_t)root_from_proof[p] - b.proof_base;
    uint32_t so = 0xffffffffu;
    if (status_out[d] == kStOk) so = account_storage_root_off(node_bytes + value_off_out[d], value_len_out[d]);
    if (so == 0xffffffffu) status = kStDepFailed;
    else rp = node_bytes + value_off_out[d] + so;
  }

  // ---- lane j owns node j of the proof (first G nodes): digest, length, offset, K2a record
  // The lane's digest and its prefetched link live in shared memory, one column per thread ([word][thread]: no bank
  // conflicts): 16 registers fewer for a kernel that otherwise needs ~146 and was capped at 64 by its occupancy target
  // (700 bytes of spills before).  A link another lane prefetched is read straight from its column.
  __shared__ uint32_t s_dg[8][kWalkThreads];
  __shared__ uint32_t s_sl[8][kWalkThreads];
  __shared__ uint32_t s_rec[4][kWalkThreads];  // the lane's node: length, K2a record, offset (lo, hi)
  uint32_t (*dg)[kWalkThreads] = reinterpret_cast<uint32_t (*)[kWalkThreads]>(&s_dg[0][threadIdx.x]);
  uint32_t (*sl)[kWalkThreads] = reinterpret_cast<uint32_t (*)[kWalkThreads]>(&s_sl[0][threadIdx.x]);
#define DG(k) (dg[k][0])
#define SL(k) (sl[k][0])
  __syncwarp(g.gmask);  // the group is done with the previous proof's columns
  uint32_t mylen = 0, mymeta = 0;
  uint64_t myoff = 0;
  bool my_is_root = false;  // digest == root: admitted to DB2 whatever its length (R5)
  if (g.lig < n) {
    const uint4* dp = reinterpret_cast<const uint4*>(digests + 32ull * (a + g.lig));
    uint4 x = MPTV_LDG(dp), y = MPTV_LDG(dp + 1);
    DG(0) = x.x; DG(1) = x.y; DG(2) = x.z; DG(3) = x.w; DG(4) = y.x; DG(5) = y.y; DG(6) = y.z; DG(7) = y.w;
    mylen = node_len[a + g.lig];
    mymeta = meta[a + g.lig];
    myoff = node_off[a + g.lig];
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) DG(i) = 0;
  }
  const bool my_short = mylen < 32;  // not admitted to DB2 unless it is the root (R5, R9)
  s_rec[0][threadIdx.x] = mylen; s_rec[1][threadIdx.x] = mymeta;
  s_rec[2][threadIdx.x] = (uint32_t)myoff; s_rec[3][threadIdx.x] = (uint32_t)(myoff >> 32);
  const uint32_t col0 = threadIdx.x & ~(uint32_t)(G - 1);  // the group's first column
  // node j's record: from its lane's column for the first G nodes of the proof (no shuffles), from memory beyond
  auto meta_of = [&](uint32_t j) -> uint32_t { return j < (uint32_t)G ? s_rec[1][col0 + j] : meta[a + j]; };
  auto len_of = [&](uint32_t j) -> uint32_t { return j < (uint32_t)G ? s_rec[0][col0 + j] : node_len[a + j]; };
  auto off_of = [&](uint32_t j) -> uint64_t {
    if (j < (uint32_t)G) return ((uint64_t)s_rec[3][col0 + j] << 32) | s_rec[2][col0 + j];
    return node_off[a + j];
  };
  // 32 bytes at q (unaligned): lanes 0..7 assemble one word each from two aligned words,
  // shuffle-broadcast to the group
  auto load_link = [&](const uint8_t* q, uint32_t (&h)[8]) {
    uint32_t w = 0;
    if (g.lig < 8) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3) + g.lig;
      const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3);
      w = __funnelshift_r(MPTV_LDG(wp), sh ? MPTV_LDG(wp + 1) : 0u, sh);  // aligned: never read past the 32 bytes
    }
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = g.bcast(w, i);
  };
  // Speculative link prefetch: in a root-first proof node j sits at depth j when every node above
  // it is a plain branch.  Lane j therefore fetches, up front and in parallel with all other lanes,
  // the child reference its own node would hand out for key nibble j.  The walk uses lane cur's
  // copy when it really arrives at node cur with path index cur, and reads memory otherwise, so
  // this only shortens the dependent-load chain (one HBM latency instead of one per level); the
  // reference's order-independent semantics (R6) are untouched.
  bool have_spec = false;
  if (g.lig < n && mymeta != kMetaSlow && meta_dec(mymeta) == kDecOk && meta_fast(mymeta)) {
    const uint32_t nibj = key_nibble(key, klen, g.lig);
    const uint32_t mk = meta_mask(mymeta);
    if (nibj < 16 && ((mk >> nibj) & 1u)) {
      const uint8_t* q = node_bytes + myoff + meta_hdr(mymeta) + nibj + 32u * __popc(mk & ((1u << nibj) - 1u)) + 1;
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
      const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3);
      uint32_t w[9];
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = MPTV_LDG(wp + i);
      w[8] = sh ? MPTV_LDG(wp + 8) : 0u;
#pragma unroll
      for (int i = 0; i < 8; i++) SL(i) = __funnelshift_r(w[i], w[i + 1], sh);
      have_spec = true;
    }
  }
  if (!have_spec) {
#pragma unroll
    for (int i = 0; i < 8; i++) SL(i) = 0;
  }
  __syncwarp(g.gmask);  // the columns are read by the other lanes of the group from here on
  // lowest node index whose digest equals h (MemoryDB keyed by hash).  `filtered` applies the DB2
  // admission rule of verify_proof: digest == root or len >= 32 (R5, R9).
  auto find = [&](const uint32_t (&h)[8], bool filtered) -> int {
    bool m = g.lig < n;
#pragma unroll
    for (int k = 0; k < 8; k++) m &= (DG(k) == h[k]);
    if (filtered && my_short && !my_is_root) m = false;
    uint32_t bal = g.ballot(m);
    if (bal) return (int)(__ffs(bal) - 1);
    for (uint32_t base = G; base < n; base += G) {  // proofs with more than G nodes (rare)
      const uint32_t i = base + g.lig;
      m = false;
      if (i < n) {
        const uint4* dp = reinterpret_cast<const uint4*>(digests + 32ull * (a + i));
        uint4 x = MPTV_LDG(dp), y = MPTV_LDG(dp + 1);
        m = x.x == h[0] && x.y == h[1] && x.z == h[2] && x.w == h[3] && y.x == h[4] && y.y == h[5] &&
            y.z == h[6] && y.w == h[7];
        if (m && filtered && node_len[a + i] < 32) {
          bool is_root = true;
          for (int k = 0; k < 8; k++) is_root &= (load_u32_unaligned(rp + 4 * k) == h[k]);
          m = is_root;
        }
      }
      bal = g.ballot(m);
      if (bal) return (int)(base + __ffs(bal) - 1);
    }
    return -1;
  };

  uint32_t cur = 0;
  if (status == kStOk) {
    // ---- lib.rs:14  EthTrie::from: root must be present (R2) and decodable (R3)
    uint32_t h[8];
    load_link(rp, h);
    {
      bool e = g.lig < n;
#pragma unroll
      for (int k = 0; k < 8; k++) e &= (DG(k) == h[k]);
      my_is_root = e;
    }
    int ri = find(h, false);
    if (ri < 0) status = kStInvalidStateRoot;
    else {
      const uint32_t m = meta_of((uint32_t)ri);
      if (meta_dec(m) == kDecErr) status = kStInvalidStateRoot;
      else if (meta_dec(m) == kDecPanic) status = kStPanicOther;
      else if (meta_kind(m) == kKindHash) {
        // commit() returns the inner hash; recover_from_db(inner) must find a decodable node or
        // root_hash() panics; if it does the assert fails because inner != root
        load_link(node_bytes + off_of((uint32_t)ri) + meta_hdr(m), h);
        int j = find(h, false);
        status = (j >= 0 && meta_dec(meta_of((uint32_t)j)) == kDecOk) ? kStRootNotCanonical : kStPanicOther;
      } else if (!meta_canon(m)) status = kStRootNotCanonical;  // lib.rs:19 (R4)
      cur = (uint32_t)ri;
    }
  }

  if (status == kStOk) {
    // ---- lib.rs:20  verify_proof -> get_at (R11-R16)
    uint32_t idx = 0;  // path index
    uint32_t lp = 0;   // offset of the current list node inside node `cur` (0 = the node itself)
    uint32_t m = meta_of(cur);
    if (meta_kind(m) == kKindEmpty) status = kStKeyNotFound;
    bool done = status != kStOk;
    for (uint32_t guard = 0; !done; guard++) {
      if (guard > 2 * klen + n + 8) { status = kStInvalidProof; break; }  // every step consumes a nibble or a node
      const uint8_t* link = nullptr;  // where the 32-byte child reference to follow lives
      bool use_spec = false;          // ... or take it from lane cur's prefetched copy
      if (lp == 0 && meta_fast(m)) {
        // plain branch: 16 children that are each 0x80 or a 32-byte hash, empty value.  Child
        // selection is a popcount over the occupancy map K2a recorded -- no node bytes are read
        // except the link itself.
        const uint32_t nib = key_nibble(key, klen, idx);
        const uint32_t mk = meta_mask(m);
        if (nib == 16 || !((mk >> nib) & 1u)) { status = kStKeyNotFound; break; }  // R14 (empty value) / R15
        const bool spec = cur < (uint32_t)G && idx == cur;  // lane cur prefetched exactly this link
        idx += 1;
        if (spec) use_spec = true;
        else link = node_bytes + off_of(cur) + meta_hdr(m) + nib + 32u * __popc(mk & ((1u << nib) - 1u)) + 1;
      } else {
        // general node (leaf, extension, branch with inline children or a value, inline node).
        // K2a validated the whole node, so the headers below are known to be well formed.
        const uint8_t* nb = node_bytes + off_of(cur);
        const uint32_t nl = len_of(cur);
        Hdr lh;
        rlp_hdr(nb + lp, nl - lp, lh);
        int cnt;
        if (lp == 0) cnt = meta_kind(m) == kKindBranch ? 17 : 2;
        else cnt = scan_items(nb, lp, lh);
        uint32_t child = 0;
        Hdr ch;
        if (cnt == 2) {
          Hdr ph;
          const uint32_t it0 = lp + lh.hdr_len;
          rlp_hdr(nb + it0, nl - it0, ph);
          const uint8_t* pp = nb + it0 + ph.hdr_len;
          const uint32_t b0 = ldb(pp);
          const uint32_t odd = (b0 >> 4) & 1, leaf = (b0 >> 5) & 1;
          const uint32_t nn = (ph.payload_len - 1) * 2 + odd;
          const uint32_t rem = 2 * klen - idx;  // key nibbles left (terminator excluded)
          bool ok = leaf ? (nn == rem) : (nn <= rem);
          if (ok) {
            // nibble-parallel compare: lane t checks nibbles t, t+G, ...
            bool eq = true;
            for (uint32_t t = g.lig; t < nn; t += G) {
              const uint32_t qn = t + 2 - odd;  // nibble position in the hex-prefix byte string
              const uint32_t pb = ldb(pp + (qn >> 1));
              const uint32_t pnib = (qn & 1) ? (pb & 15) : (pb >> 4);
              eq &= (pnib == key_nibble(key, klen, idx + t));
            }
            ok = g.all(eq);
          }
          const uint32_t it1 = it0 + ph.hdr_len + ph.payload_len;
          if (!ok) { status = kStKeyNotFound; break; }
          rlp_hdr(nb + it1, nl - it1, ch);
          if (leaf) {
            if (ch.payload_len == 1) { voff = it1; vlen = ch.hdr_len + 1; }  // R20
            else { voff = it1 + ch.hdr_len; vlen = ch.payload_len; }
            break;
          }
          idx += nn;
          child = it1;
        } else {
          const uint32_t nib = key_nibble(key, klen, idx);
          uint32_t q = lp + lh.hdr_len;
          for (uint32_t i = 0; i < nib; i++) {  // nib == 16 walks to the value item
            Hdr t;
            rlp_hdr(nb + q, nl - q, t);
            q += t.hdr_len + t.payload_len;
          }
          rlp_hdr(nb + q, nl - q, ch);
          if (nib == 16) {
            // branch value (R14); empty => None
            uint32_t vo, vl;
            if (ch.payload_len == 1) { vo = q; vl = ch.hdr_len + 1; }
            else { vo = q + ch.hdr_len; vl = ch.payload_len; }
            if (vl == 0) status = kStKeyNotFound;
            else { voff = vo; vlen = vl; }
            break;
          }
          idx += 1;
          child = q;
        }
        if (ch.is_list) { lp = child; continue; }                         // inline node (R16)
        if (ch.payload_len == 0) { status = kStKeyNotFound; break; }      // empty slot (R15)
        link = nb + child + ch.hdr_len;
      }
      // ---- follow a hash reference: shuffle-broadcast the link, compare against every digest
      uint32_t h[8];
      if (use_spec) {
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = s_sl[i][col0 + cur];  // lane cur's column
      } else {
        load_link(link, h);
      }
      for (uint32_t hops = 0;; hops++) {
        const int j = find(h, true);
        if (j < 0 || hops > n) { status = kStInvalidProof; done = true; break; }  // R8 / R9
        const uint32_t mj = meta_of((uint32_t)j);
        if (meta_dec(mj) == kDecErr) { status = kStInvalidProof; done = true; break; }
        if (meta_dec(mj) == kDecPanic) { status = kStPanicOther; done = true; break; }
        if (meta_kind(mj) == kKindEmpty) { status = kStKeyNotFound; done = true; break; }
        cur = (uint32_t)j; m = mj; lp = 0;
        if (meta_kind(mj) != kKindHash) break;
        load_link(node_bytes + off_of(cur) + meta_hdr(mj), h);  // a node that is itself a bare hash reference
      }
    }
  }

  const uint64_t base_off = off_of(cur);
  if (g.lig == 0) {
    status_out[p] = (uint8_t)status;
    const bool okv = status == kStOk;
    value_off_out[p] = okv ? base_off + voff : 0ull;
    value_len_out[p] = okv ? vlen : 0u;
  }
#undef DG
#undef SL
}

// ------------------------------------------------------------------ K2f: one thread per proof, the common case
// A well-formed proof as eth_trie's get_proof emits it is a CHAIN: node 0 hashes to the root, node i is
// a plain branch (16 x {empty | 32-byte hash}, no value) whose child for the key's nibble i is the
// hash of node i+1, and the chain ends in a plain leaf or an empty slot.  For such a proof the
// reference's hash-keyed lookups (R5-R8) can only ever find node i+1 (or a byte-identical copy), so
// the verdict and value follow from comparing each link with the NEXT digest -- one thread, no
// shuffles, ~40 instructions per level.  Anything else (shuffled or junk-interleaved order,
// extensions, inline children, branch values, short nodes, non-canonical roots, decode errors,
// missing nodes, tampered bytes) is not judged here: the proof is appended to the deferred list and
// the cooperative kernel K2b, which implements the full rule set, decides it.
__device__ __forceinline__ bool eq32_unaligned(const uint8_t* q, const uint8_t* d32 /* 16-byte aligned */) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
  const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3);
  const uint4 x = MPTV_LDG(reinterpret_cast<const uint4*>(d32)), y = MPTV_LDG(reinterpret_cast<const uint4*>(d32) + 1);
  uint32_t w[9];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = MPTV_LDG(wp + i);
  w[8] = sh ? MPTV_LDG(wp + 8) : 0u;  // an aligned reference ends exactly at its 8th word: never read past it
  const uint32_t d[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
  uint32_t diff = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) diff |= __funnelshift_r(w[i], w[i + 1], sh) ^ d[i];
  return diff == 0;
}

// does any of the proof's n digests equal the 32 bytes at q?  (first word first: a miss costs one load per node)
__device__ bool digest_present(const uint8_t* q, const uint8_t* dg, uint32_t n) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
  const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3);
  const uint32_t w0 = __funnelshift_r(MPTV_LDG(wp), sh ? MPTV_LDG(wp + 1) : 0u, sh);
  for (uint32_t t = 0; t < n; t++)
    if (MPTV_LDG(reinterpret_cast<const uint32_t*>(dg + 32ull * t)) == w0 && eq32_unaligned(q, dg + 32ull * t)) return true;
  return false;
}

constexpr int kFastPre = 10;  // levels whose records and links K2f loads up front (a state proof has 6 ... 10 nodes)

// true: decided (outputs written); false: defer to K2b
__device__ bool fast_one(const DeviceBatch& b, uint64_t p, bool dependent, const uint8_t* __restrict__ digests,
                         const uint32_t* __restrict__ meta, uint8_t* status_out, uint64_t* value_off_out,
                         uint32_t* value_len_out) {
  const uint8_t* __restrict__ node_bytes = b.node_bytes - b.byte_base;
  const uint32_t a = b.proof_first[p] - b.node_base, n = b.proof_first[p + 1] - b.proof_first[p];
  if (n == 0) return false;
  const uint8_t* key = b.key_bytes + (b.key_off[p] - b.key_base);
  const uint32_t klen = b.key_len ? b.key_len[p] : b.key_off[p + 1] - b.key_off[p];
  const uint8_t* rp = b.roots + 32 * p;
  if (dependent) {
    // storage-circuit main.rs:10-27: the root is the storage_root of the account the earlier proof returned
    const uint64_t d = (uint64_t)b.root_from_proof[p] - b.proof_base;
    uint32_t so = 0xffffffffu;
    if (status_out[d] == kStOk) so = account_storage_root_off(node_bytes + value_off_out[d], value_len_out[d]);
    if (so == 0xffffffffu) {  // account proof rejected, or its value is not an Account RLP
      status_out[p] = (uint8_t)kStDepFailed; value_off_out[p] = 0; value_len_out[p] = 0;
      return true;
    }
    rp = node_bytes + value_off_out[d] + so;
  }
  if (!eq32_unaligned(rp, digests + 32ull * a)) {
    // node 0 is not the root.  If NO supplied node hashes to the root the verdict is InvalidStateRoot whatever
    // the order (R2: tampered / dropped root node, wrong root); if some other node does, K2b sorts it out.
    if (digest_present(rp, digests + 32ull * a, n)) return false;
    status_out[p] = (uint8_t)kStInvalidStateRoot; value_off_out[p] = 0; value_len_out[p] = 0;
    return true;
  }
  // The records and offsets of the proof's first kFastPre nodes, loaded together; then every level's link and the
  // digest it should equal, ALL issued before the first decision.  Node i is a plain branch at path index i as long
  // as every node above it is one -- the chain shape this kernel checks -- so no level's addresses depend on another
  // level's outcome, and the thread has the whole proof's loads in flight at once instead of one dependent round
  // trip per level.  (ncu, config 2: the level-at-a-time loop re-fetched the record sectors at every level -- 6.4 L1
  // sector misses a level for 4.4 algorithmic -- and ran at 53 % of HBM bandwidth waiting on itself.)
  uint32_t mm[kFastPre];
  uint64_t oo[kFastPre];
#pragma unroll
  for (int i = 0; i < kFastPre; i++) {
    mm[i] = (uint32_t)i < n ? MPTV_LDG(meta + a + i) : kMetaSlow;
    oo[i] = (uint32_t)i < n ? MPTV_LDG(b.node_off + a + i) : 0ull;
  }
  uint32_t eq = 0;  // bit i: node i's link for the key's nibble i equals the digest of node i + 1
#pragma unroll
  for (int i = 0; i < kFastPre - 1; i++) {
    const uint32_t m = mm[i];
    if ((uint32_t)(i + 1) < n && m != kMetaSlow && meta_dec(m) == kDecOk && meta_kind(m) == kKindBranch && meta_fast(m)) {
      const uint32_t nib = key_nibble(key, klen, (uint32_t)i);
      const uint32_t mk = meta_mask(m);
      if (nib < 16 && ((mk >> nib) & 1u)) {
        const uint8_t* link = node_bytes + oo[i] + meta_hdr(m) + nib + 32u * __popc(mk & ((1u << nib) - 1u)) + 1;
        eq |= (uint32_t)eq32_unaligned(link, digests + 32ull * (a + i + 1)) << i;
      }
    }
  }
  // one level of the reference's walk: 0 = go on, 1 = decided (outputs written), 2 = not chain-shaped (K2b decides)
  auto level = [&](uint32_t i, uint32_t m, uint64_t off, bool have_eq, bool link_ok) -> int {
    if (m == kMetaSlow || meta_dec(m) != kDecOk) return 2;
    const bool plain_branch = meta_kind(m) == kKindBranch && meta_fast(m);
    // a plain branch's length follows from its record (header + 16 + 1 item bytes + 32 per hashed child):
    // no need to touch node_len for the levels above the leaf
    const uint32_t nl = plain_branch ? meta_hdr(m) + 17u + 32u * __popc(meta_mask(m)) : b.node_len[a + i];
    if (i == 0 ? !meta_canon(m) : nl < 32) return 2;  // lib.rs:19 on the root; R9 admission otherwise
    if (plain_branch) {
      const uint32_t nib = key_nibble(key, klen, i);  // the path index equals the level: every node above is a plain branch
      const uint32_t mk = meta_mask(m);
      if (nib == 16 || !((mk >> nib) & 1u)) {  // R14 (no value in a plain branch) / R15
        status_out[p] = (uint8_t)kStKeyNotFound; value_off_out[p] = 0; value_len_out[p] = 0;
        return 1;
      }
      const uint8_t* link = node_bytes + off + meta_hdr(m) + nib + 32u * __popc(mk & ((1u << nib) - 1u)) + 1;
      if (i + 1 < n && !have_eq) link_ok = eq32_unaligned(link, digests + 32ull * (a + i + 1));
      if (i + 1 >= n || !link_ok) {
        // the next node is not the child.  If NO supplied node has that digest the reference's lookup misses
        // (R8: truncated proof, tampered child, wrong key into an unproven subtree) -> InvalidProof; if some
        // node does (shuffled or padded proofs) the full rule set of K2b decides.
        if (digest_present(link, digests + 32ull * a, n)) return 2;
        status_out[p] = (uint8_t)kStInvalidProof; value_off_out[p] = 0; value_len_out[p] = 0;
        return 1;
      }
      return 0;
    }
    if (meta_kind(m) != kKindLeaf) return 2;
    const uint8_t* nb = node_bytes + off;
    Hdr lh, ph, ch;
    rlp_hdr(nb, nl, lh);  // validated by K1 / K2a
    const uint32_t it0 = lh.hdr_len;
    rlp_hdr(nb + it0, nl - it0, ph);
    const uint8_t* pp = nb + it0 + ph.hdr_len;
    const uint32_t odd = (ldb(pp) >> 4) & 1;
    const uint32_t nn = (ph.payload_len - 1) * 2 + odd;
    bool ok = nn == 2 * klen - i;  // R12: the whole remaining path, length and nibbles
    for (uint32_t t = 0; ok && t < nn; t++) {
      const uint32_t qn = t + 2 - odd;
      const uint32_t pb = ldb(pp + (qn >> 1));
      ok = ((qn & 1) ? (pb & 15) : (pb >> 4)) == key_nibble(key, klen, i + t);
    }
    if (!ok) { status_out[p] = (uint8_t)kStKeyNotFound; value_off_out[p] = 0; value_len_out[p] = 0; return 1; }
    const uint32_t it1 = it0 + ph.hdr_len + ph.payload_len;
    rlp_hdr(nb + it1, nl - it1, ch);
    status_out[p] = (uint8_t)kStOk;
    if (ch.payload_len == 1) { value_off_out[p] = off + it1; value_len_out[p] = ch.hdr_len + 1; }  // R20
    else { value_off_out[p] = off + it1 + ch.hdr_len; value_len_out[p] = ch.payload_len; }
    return 1;
  };
#pragma unroll
  for (int i = 0; i < kFastPre; i++) {
    if ((uint32_t)i >= n) return false;
    const int r = level((uint32_t)i, mm[i], oo[i], i < kFastPre - 1, (eq >> i) & 1u);
    if (r) return r == 1;
  }
  for (uint32_t i = kFastPre; i < n; i++) {  // proofs longer than the preloaded window (rare): level at a time
    const int r = level(i, meta[a + i], b.node_off[a + i], false, false);
    if (r) return r == 1;
  }
  return false;
}


}  // namespace
}  // namespace mptv
