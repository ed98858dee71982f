// keccak_kernels.cu -- K0 (rate-block binning), K1 (batched Keccak-256 of proof nodes) and K1L (the trie
// rebuild's fused leaf encode + Keccak-256, further down).
//
// K1 replaces crypto_ops::keccak::digest_keccak (/root/reference/crypto-ops/src/keccak.rs:6-12)
// applied to every proof node (/root/reference/crypto-ops/src/lib.rs:10-13) for a whole batch:
// one thread per node, state in registers (keccak_f1600.cuh), nodes read from a flat CSR arena.
// Nodes are binned by their 136-byte rate-block count so that the 32 lanes of a warp run the
// same number of permutations.  Node bytes are staged HBM -> shared memory one rate block at a
// time by 16-byte asynchronous copies (cp.async / LDGSTS) into a 2-stage per-thread ring, so the
// copy of blocks k+1, k+2 overlaps the permutation of block k.  (A first version issued one 1-D
// bulk copy (cp.async.bulk / UBLKCP) per thread per block: UBLKCP takes uniform-register
// operands, so the compiler serialised it over the 32 lanes -- ELECT + 5 R2UR + UBLKCP per lane,
// 7.6 % of all issued instructions in profiles/r01_keccak_v1_ncu.txt -- and it was replaced.)
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "keccak_f1600.cuh"
#include "kernels.h"
#include "trie_rec.cuh"

namespace mptv {

// ------------------------------------------------------------------ K0: binning
// bin(nb): exact for nb <= 64, then 16-block-wide bins, last bin is a catch-all.
__host__ __device__ __forceinline__ uint32_t nb_bin(uint32_t nb) {
  if (nb <= 64) return nb;
  uint32_t b = 64 + (nb - 64 + 15) / 16;
  return b < kNumBins - 1 ? b : kNumBins - 1;
}

// ids (may be NULL = identity): the list of node indices to bin -- the rebuild path bins one trie
// level at a time out of a larger node table.
// keep (may be NULL): only nodes with keep[i] == i take part (dedup_nodes: the representatives); then
// totals[0] / totals[1] receive the number of kept nodes and their Keccak-f count.
__global__ void __launch_bounds__(256) k_bin_hist(const uint32_t* __restrict__ node_len,
                                                  const uint32_t* __restrict__ ids, uint64_t n_nodes,
                                                  uint32_t* __restrict__ hist, const uint32_t* __restrict__ keep,
                                                  unsigned long long* __restrict__ totals) {
  __shared__ uint32_t sh[kNumBins];
  __shared__ unsigned long long s_perm;
  for (int i = threadIdx.x; i < kNumBins; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) s_perm = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kBinNodesPerBlock;
  uint64_t end = base + kBinNodesPerBlock < n_nodes ? base + kBinNodesPerBlock : n_nodes;
  uint32_t perms = 0;
  for (uint64_t i = base + threadIdx.x; i < end; i += blockDim.x) {
    const uint32_t id = ids ? ids[i] : (uint32_t)i;
    if (keep && keep[id] != id) continue;
    const uint32_t nb = node_len[id] / 136 + 1;
    perms += nb;
    atomicAdd(&sh[nb_bin(nb)], 1u);
  }
  if (totals && perms) atomicAdd(&s_perm, (unsigned long long)perms);
  __syncthreads();
  uint32_t kept = 0;
  for (int i = threadIdx.x; i < kNumBins; i += blockDim.x)
    if (sh[i]) { atomicAdd(&hist[i], sh[i]); kept += sh[i]; }
  if (totals) {
    if (kept) atomicAdd(&totals[0], (unsigned long long)kept);
    if (threadIdx.x == 0 && s_perm) atomicAdd(&totals[1], s_perm);
  }
}

// bins laid out in DESCENDING block count so the long nodes start first (tail balance)
// cursor[kNumBins + 1] (the word after K1's tile counter) receives the number of nodes in bins >= long_bin:
// order[0 .. split) are the "long" nodes that K1 / K1L hash in a launch of their own
__global__ void k_bin_scan(const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, int long_bin) {
  if (threadIdx.x == 0) {
    uint32_t acc = 0, split = 0;
    for (int b = kNumBins - 1; b >= 0; b--) {
      cursor[b] = acc;
      acc += hist[b];
      if (b == long_bin) split = acc;
    }
    cursor[kNumBins + 1] = split;
  }
}

__global__ void __launch_bounds__(256) k_bin_scatter(const uint32_t* __restrict__ node_len,
                                                     const uint32_t* __restrict__ ids, uint64_t n_nodes,
                                                     uint32_t* __restrict__ cursor,
                                                     uint32_t* __restrict__ order, const uint32_t* __restrict__ keep) {
  __shared__ uint32_t cnt[kNumBins];
  __shared__ uint32_t start[kNumBins];
  for (int i = threadIdx.x; i < kNumBins; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kBinNodesPerBlock;
  uint64_t end = base + kBinNodesPerBlock < n_nodes ? base + kBinNodesPerBlock : n_nodes;
  constexpr int kPer = kBinNodesPerBlock / 256;
  uint32_t br[kPer];  // bin << 16 | rank within this CTA's share of the bin; 0xffffffff = not kept
  uint32_t id[kPer];
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    uint64_t i = base + threadIdx.x + (uint64_t)j * 256;
    br[j] = 0xffffffffu;
    if (i < end) {
      id[j] = ids ? ids[i] : (uint32_t)i;
      if (keep && keep[id[j]] != id[j]) continue;
      const uint32_t b = nb_bin(node_len[id[j]] / 136 + 1);
      br[j] = (b << 16) | atomicAdd(&cnt[b], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumBins; i += blockDim.x)
    if (cnt[i]) start[i] = atomicAdd(&cursor[i], cnt[i]);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    uint64_t i = base + threadIdx.x + (uint64_t)j * 256;
    if (i < end && br[j] != 0xffffffffu) order[start[br[j] >> 16] + (br[j] & 0xffffu)] = id[j];
  }
}

// ------------------------------------------------------------------ K1: Keccak-256
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte asynchronous copy global -> shared (LDGSTS), L2-only caching: each lane streams its own
// node, so the bytes are used exactly once by this SM.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// Stage geometry: rate block k of a node covers node bytes [136k, 136k+136).  With the node
// 16-byte aligned in the arena the 16-byte aligned window that contains it is
// [136k - 8*(k&1), +144) = nine 16-byte chunks.  One 144-byte slot per thread per stage; the slot
// stride of 144 B = 9 x 16 B keeps the per-lane LDS.64 / LDGSTS.128 accesses spread over the banks.
// A thread only ever touches its own slots, and cp.async completion (wait_group) is tracked per
// thread, so the ring needs no barrier at all.
constexpr int kSlotBytes = 144;
constexpr int kStages = 2;

// issue the copies of block k of a node of `len` bytes into the slot at `dst` (one commit group)
__device__ __forceinline__ void stage_block(uint32_t dst, const uint8_t* src, uint32_t k, uint32_t len, bool on) {
  if (on) {
    const uint32_t a = 136u * k - 8u * (k & 1u);
    uint32_t e = 136u * k + 136u;
    if (e > len) e = len;
    const uint8_t* s = src + a;
#pragma unroll
    for (int c = 0; c < 9; c++)
      if (a + 16u * c < e) cp_async16(dst + 16 * c, s + 16 * c);
  }
  cp_async_commit();
}


// ---- fused K2a fast path ---------------------------------------------------------------------
// While block k of a node sits in this thread's shared-memory slot, classify the node without a
// second pass over HBM.  Only the two shapes that make up real proofs are decided here:
//   * plain branch: list of 17 items, each child 0x80 or 0xa0 + 32 bytes, value 0x80
//   * plain leaf:   list of [hex-prefix string with flag 2/3 and canonical pad, string value that is
//                   not the 2-byte form 0x81 b], nothing trailing
// Everything else (extensions, inline children, branch values, odd RLP, trailing bytes, lists
// >= 64 KiB) is marked kMetaSlow and decided by k_parse_nodes, which implements the full rule
// set.  For the shapes decided here the record equals what k_parse_nodes would produce.
constexpr uint32_t kClsScan = 0, kClsDone = 1, kClsOff = 2;

__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// p = shared address of node byte 136*k.  cmeta carries the list header length between calls.
__device__ __forceinline__ void classify_block(uint32_t p, uint32_t k, uint32_t len, uint32_t& pos, uint32_t& cm,
                                               uint32_t& cls, uint32_t& cmeta) {
  const uint32_t blk0 = 136u * k;
  if (k == 0) {
    cls = kClsDone;  // pessimistic: cmeta == kMetaSlow unless a shape below matches
    if (len < 2) return;
    const uint32_t b0 = lds8(p);
    uint32_t hl, pay;
    if (b0 >= 0xc0 && b0 <= 0xf7) { hl = 1; pay = b0 - 0xc0; }
    else if (b0 == 0xf8) { hl = 2; pay = lds8(p + 1); if (pay < 56) return; }
    else if (b0 == 0xf9 && len >= 3) { hl = 3; pay = (lds8(p + 1) << 8) | lds8(p + 2); if (pay < 256) return; }
    else return;
    if (hl + pay != len || pay == 0) return;
    const uint32_t b = lds8(p + hl);
    if (b != 0x80 && b != 0xa0) {
      // candidate leaf: [path, value]
      uint32_t phl, ppl;
      if (b < 0x80) { phl = 0; ppl = 1; }
      else if (b <= 0xb7) { phl = 1; ppl = b - 0x80; }
      else return;
      const uint32_t q = hl + phl + ppl;  // value item
      if (ppl < 1 || q >= len) return;
      const uint32_t f = lds8(p + hl + phl);
      const uint32_t flag = f >> 4;
      if (!(flag == 3 || (flag == 2 && (f & 15) == 0))) return;  // leaf with canonical pad only
      if (ppl == 1 && phl == 1) return;                          // 0x81 xx would be non-canonical RLP here
      const uint32_t v = lds8(p + q);
      uint32_t vhl, vpl;
      if (v < 0x80) { vhl = 0; vpl = 1; }
      else if (v <= 0xb7) { vhl = 1; vpl = v - 0x80; if (vpl == 1) return; }
      else if (v == 0xb8) { if (q + 1 >= len) return; vhl = 2; vpl = lds8(p + q + 1); if (vpl < 56) return; }
      else if (v == 0xb9) { if (q + 2 >= len) return; vhl = 3; vpl = (lds8(p + q + 1) << 8) | lds8(p + q + 2); if (vpl < 256) return; }
      else return;
      if (q + vhl + vpl != len) return;
      cmeta = make_meta(kKindLeaf, kDecOk, 1, 0, hl, 0);
      return;
    }
    cls = kClsScan;
    pos = hl;
    cmeta = hl;  // remembered for the final record
  }
  uint32_t end = blk0 + 136u;
  if (end > len) end = len;
  while (pos < end) {
    const uint32_t b = lds8(p + (pos - blk0));
    const uint32_t cnt = cm >> 20;
    if (b == 0xa0) { cm |= 1u << cnt; pos += 33; }
    else if (b == 0x80) pos += 1;
    else { cls = kClsDone; cmeta = kMetaSlow; return; }
    cm += 1u << 20;
    if (cnt >= 17) { cls = kClsDone; cmeta = kMetaSlow; return; }
  }
}

__global__ void __launch_bounds__(kKeccakThreads, kKeccakMinBlocks)
k_keccak256_nodes(const uint8_t* __restrict__ node_bytes, uint64_t byte_base, const uint64_t* __restrict__ node_off,
                  const uint32_t* __restrict__ node_len, const uint32_t* __restrict__ order,
                  uint64_t n_nodes_all, uint8_t* __restrict__ digests, uint32_t* __restrict__ meta_out,
                  uint32_t* __restrict__ tile_counter, const uint32_t* __restrict__ split, int part) {
  extern __shared__ __align__(128) uint8_t smem[];  // [stage][thread] slots
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x;
  const uint32_t slot0 = smem_u32(smem + tid * kSlotBytes);
  constexpr uint32_t kStageStride = kKeccakThreads * kSlotBytes;

  // Tiles of 128 nodes are handed out by an atomic counter: with the nodes sorted by DESCENDING block
  // count this is longest-processing-time-first scheduling, so the CTAs finish together (a static
  // round-robin leaves the CTA that draws the longest tile of every round far behind on long-tailed
  // inputs such as receipt tries).
  // split != NULL: this launch hashes one side of order[] -- part 0 the long nodes [0, *split), part 1 the rest.
  // A long sequential chain (a 30 KB receipt leaf is 221 permutations) that shares a scheduler with warps of
  // short nodes is starved until everything else has drained, so the long side runs first, on its own, with one
  // warp per scheduler (Keccak's instruction-level parallelism keeps the alu pipe busy from a single warp).
  const uint64_t n_long = split ? (uint64_t)*split : 0;
  const uint64_t first = (split && part == 1) ? n_long : 0;
  const uint64_t n_nodes = split ? (part == 0 ? n_long : n_nodes_all - n_long) : n_nodes_all;
  const uint64_t n_tiles = (n_nodes + kKeccakThreads - 1) / kKeccakThreads;
  for (uint64_t tile = blockIdx.x;; tile += gridDim.x) {
    if (tile_counter) {
      __syncthreads();
      if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
      __syncthreads();
      tile = s_tile;
    }
    if (tile >= n_tiles) break;
    const uint64_t slot_idx = first + tile * kKeccakThreads + tid;
    const bool have = slot_idx < first + n_nodes;
    uint32_t node = 0, len = 0, nb = 0;
    const uint8_t* src = node_bytes;
    if (have) {
      node = order ? order[slot_idx] : (uint32_t)slot_idx;
      len = node_len[node];
      nb = len / 136u + 1u;
      src = node_bytes + (node_off[node] - byte_base);
    }
    uint32_t lo[25], hi[25];
#pragma unroll
    for (int i = 0; i < 25; i++) { lo[i] = 0; hi[i] = 0; }
    // fused K2a fast path (see classify_* below): scan position, item count << 20 | occupancy map
    uint32_t cpos = 0, ccm = 0, ccls = meta_out ? kClsScan : kClsOff, cmeta = kMetaSlow;

    // prologue: fill the ring (empty commit groups keep the group count uniform)
#pragma unroll
    for (int s = 0; s < kStages; s++) stage_block(slot0 + s * kStageStride, src, s, len, (uint32_t)s < nb);

    for (uint32_t k = 0; k < nb; k++) {
      const uint32_t s = k % kStages;
      cp_async_wait<kStages - 1>();  // block k has landed in this thread's slot
      const uint32_t p = slot0 + s * kStageStride + 8u * (k & 1u);
      const uint32_t valid = len - 136u * k;  // message bytes from the start of this block
      if (ccls == kClsScan) classify_block(p, k, len, cpos, ccm, ccls, cmeta);
      if (valid >= 136u) {
#pragma unroll
        for (int j = 0; j < 17; j++) {
          uint2 w = lds64(p + 8 * j);
          lo[j] ^= w.x;
          hi[j] ^= w.y;
        }
      } else {
        // last block: message tail, then pad10*1 with the Keccak delimiter 0x01 (keccak.rs:7 v256).
        // Branch-free per 32-bit word; stale slot bytes beyond the copied range are masked off.
#pragma unroll
        for (int j = 0; j < 17; j++) {
          uint2 w = lds64(p + 8 * j);
          uint32_t v[2] = {w.x, w.y};
#pragma unroll
          for (int hlf = 0; hlf < 2; hlf++) {
            const int wi = 2 * j + hlf;
            const int keep = (int)valid - 4 * wi;  // message bytes in this word
            const uint32_t msk = keep >= 4 ? 0xffffffffu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
            uint32_t x = v[hlf] & msk;
            if ((int)(valid >> 2) == wi) x ^= 1u << (8u * (valid & 3u));
            v[hlf] = x;
          }
          lo[j] ^= v[0];
          hi[j] ^= v[1];
        }
        hi[16] ^= 0x80000000u;
      }
      // the slot now lives in registers: refill it with block k + kStages while we permute
      stage_block(slot0 + s * kStageStride, src, k + kStages, len, k + kStages < nb);
      keccak_f1600(lo, hi);
    }
    cp_async_wait<0>();
    if (have && meta_out) {
      if (ccls == kClsScan) {
        // every block scanned without leaving the {0x80, 0xa0+32} alphabet: a plain branch iff it
        // has exactly 17 items, ends exactly at the node's end and its value item is empty
        const uint32_t cnt = ccm >> 20, mask = ccm & 0x1ffffu;
        if (cnt == 17 && cpos == len && !(mask >> 16)) cmeta = make_meta(kKindBranch, kDecOk, 1, 1, cmeta, mask);
        else cmeta = kMetaSlow;
      }
      meta_out[node] = cmeta;
    }
    if (have) {
      uint4* out = reinterpret_cast<uint4*>(digests + (uint64_t)node * 32u);
      out[0] = make_uint4(lo[0], hi[0], lo[1], hi[1]);
      out[1] = make_uint4(lo[2], hi[2], lo[3], hi[3]);
    }
  }
}

// ------------------------------------------------------------------ K1L: fused leaf encode + Keccak-256
// Trie rebuild, leaf level (> 95 % of the bytes of a tx / receipt trie): the leaf node
//     rlp([hex_prefix(path), value])            eth_trie write_node, Node::Leaf
// is never materialised.  Each thread builds the <= 42-byte RLP prefix (list header, path item, value
// header) of its leaf in shared memory and streams the VALUE bytes straight from the caller's value
// arena through the same per-thread cp.async ring as K1; the byte misalignment between the node and
// the 16-byte aligned value (the prefix length) is taken out with one funnel shift per absorbed word.
// This removes one write + one read of every leaf byte from HBM and a whole encode launch.
constexpr int kLeafHead = 48;                       // room in front of the window for the prefix of block 0
constexpr int kLeafSlotBytes = kLeafHead + 160;     // ten 16-byte chunks cover 136 bytes at any alignment

// issue the copies for block k: value bytes [136k - P, 136k + 136 - P) clipped to [0, vl)
__device__ __forceinline__ void stage_leaf_block(uint32_t dst, const uint8_t* val, uint32_t k, uint32_t P, uint32_t vl,
                                                 bool on) {
  if (on) {
    const uint32_t vstart = k ? 136u * k - P : 0u;
    const uint32_t a = vstart & ~15u;
    uint32_t e = 136u * k + 136u - P;
    if (e > vl) e = vl;
    const uint8_t* s = val + a;
#pragma unroll
    for (int c = 0; c < 10; c++)
      if (a + 16u * c < e) cp_async16(dst + kLeafHead + 16 * c, s + 16 * c);
  }
  cp_async_commit();
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kKeccakThreads, kKeccakMinBlocks)
k_keccak256_leaves(const TrieBatchDev in, const uint4* __restrict__ rec, const uint32_t* __restrict__ node_len,
                   const uint32_t* __restrict__ order, uint32_t n_nodes_all, uint8_t* __restrict__ digests,
                   uint32_t* __restrict__ tile_counter, const uint32_t* __restrict__ split, int part) {
  extern __shared__ __align__(128) uint8_t smem[];  // [stage][thread] slots
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x;
  const uint32_t slot0 = smem_u32(smem + tid * kLeafSlotBytes);
  constexpr uint32_t kStageStride = kKeccakThreads * kLeafSlotBytes;

  const uint32_t n_long = split ? *split : 0u;  // long / short sides of order[]: see k_keccak256_nodes
  const uint32_t first = (split && part == 1) ? n_long : 0u;
  const uint32_t n_nodes = split ? (part == 0 ? n_long : n_nodes_all - n_long) : n_nodes_all;
  const uint32_t n_tiles = (n_nodes + kKeccakThreads - 1) / kKeccakThreads;
  for (uint32_t tile = blockIdx.x;; tile += gridDim.x) {  // dynamic tiles: see k_keccak256_nodes
    if (tile_counter) {
      __syncthreads();
      if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
      __syncthreads();
      tile = s_tile;
    }
    if (tile >= n_tiles) break;
    const uint32_t slot_idx = first + tile * kKeccakThreads + tid;
    const bool have = slot_idx < first + n_nodes;
    uint32_t node = 0, len = 0, nb = 0, vl = 0, P = 0;
    const uint8_t* val = in.value_bytes;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (have) {
      node = order[slot_idx];
      r = rec[node];
      len = node_len[node];
      nb = len / 136u + 1u;
      vl = in.value_len[r.y];
      val = in.value_bytes + in.value_off[r.y];
      P = len - vl;  // list header + path item + value header (0 when the value encodes as itself)
    }
    uint32_t lo[25], hi[25];
#pragma unroll
    for (int i = 0; i < 25; i++) { lo[i] = 0; hi[i] = 0; }
#pragma unroll
    for (int s = 0; s < kStages; s++) stage_leaf_block(slot0 + s * kStageStride, val, s, P, vl, (uint32_t)s < nb);
    if (have) {
      // the prefix, right in front of value byte 0 in the stage-0 slot
      uint32_t q = slot0 + kLeafHead - P;
      const uint32_t payload = payload_of(len), hl = len - payload;
      if (hl == 1) sts8(q, 0xC0u + payload);
      else {
        sts8(q, 0xF7u + (hl - 1));
        for (uint32_t i = 1; i < hl; i++) sts8(q + i, (payload >> (8 * (hl - 1 - i))) & 0xffu);
      }
      q += hl;
      const uint32_t ps = rec_ps(r), pl = rec_pl(r), hpn = pl / 2 + 1;
      const uint8_t* key = in.key_bytes + in.key_off[r.y];
      if (hpn > 1) sts8(q++, 0x80u + hpn);
      for (uint32_t i = 0; i < hpn; i++) sts8(q + i, hp_byte(key, ps, pl, true, i));
      q += hpn;
      const uint32_t vh = slot0 + kLeafHead - q;  // bytes left for the value's string header
      if (vh == 1) sts8(q, 0x80u + vl);
      else if (vh > 1) {
        sts8(q, 0xB7u + (vh - 1));
        for (uint32_t i = 1; i < vh; i++) sts8(q + i, (vl >> (8 * (vh - 1 - i))) & 0xffu);
      }
    }
    for (uint32_t k = 0; k < nb; k++) {
      const uint32_t s = k % kStages;
      cp_async_wait<kStages - 1>();
      const uint32_t base = k ? (uint32_t)kLeafHead + ((136u * k - P) & 15u) : (uint32_t)kLeafHead - P;
      const uint32_t p = slot0 + s * kStageStride + (base & ~3u);
      const uint32_t beta = 8u * (base & 3u);
      const uint32_t valid = len - 136u * k;
      uint32_t prev = lds32(p);
      if (valid >= 136u) {
#pragma unroll
        for (int j = 0; j < 17; j++) {
          const uint32_t a = lds32(p + 8 * j + 4), b = lds32(p + 8 * j + 8);
          lo[j] ^= __funnelshift_r(prev, a, beta);
          hi[j] ^= __funnelshift_r(a, b, beta);
          prev = b;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 17; j++) {
          const uint32_t a = lds32(p + 8 * j + 4), b = lds32(p + 8 * j + 8);
          uint32_t v[2] = {__funnelshift_r(prev, a, beta), __funnelshift_r(a, b, beta)};
          prev = b;
#pragma unroll
          for (int hlf = 0; hlf < 2; hlf++) {
            const int wi = 2 * j + hlf;
            const int keep = (int)valid - 4 * wi;
            const uint32_t msk = keep >= 4 ? 0xffffffffu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
            uint32_t x = v[hlf] & msk;
            if ((int)(valid >> 2) == wi) x ^= 1u << (8u * (valid & 3u));
            v[hlf] = x;
          }
          lo[j] ^= v[0];
          hi[j] ^= v[1];
        }
        hi[16] ^= 0x80000000u;
      }
      stage_leaf_block(slot0 + s * kStageStride, val, k + kStages, P, vl, k + kStages < nb);
      keccak_f1600(lo, hi);
    }
    cp_async_wait<0>();
    if (have) {
      uint4* out = reinterpret_cast<uint4*>(digests + (uint64_t)node * 32u);
      out[0] = make_uint4(lo[0], hi[0], lo[1], hi[1]);
      out[1] = make_uint4(lo[2], hi[2], lo[3], hi[3]);
    }
  }
}

// ------------------------------------------------------------------ gather (pull mode of the streamed borsh entry)
// One warp per record: the lanes read the 16-byte aligned chunks that cover the source bytes (mapped page-locked host
// memory: one PCIe read per chunk, each chunk read once -- the neighbour's half comes by shuffle), shift them to the
// destination's alignment and store whole 16-byte chunks, the tail padded with zeros.
__device__ __forceinline__ uint4 shift_bytes(const uint4& a, const uint4& b, uint32_t sh /* 0..15, warp-uniform */) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const uint32_t bits = 8u * (sh & 3u);
  uint4 o;
  switch (sh >> 2) {
    case 0: o.x = __funnelshift_r(w[0], w[1], bits); o.y = __funnelshift_r(w[1], w[2], bits); o.z = __funnelshift_r(w[2], w[3], bits); o.w = __funnelshift_r(w[3], w[4], bits); break;
    case 1: o.x = __funnelshift_r(w[1], w[2], bits); o.y = __funnelshift_r(w[2], w[3], bits); o.z = __funnelshift_r(w[3], w[4], bits); o.w = __funnelshift_r(w[4], w[5], bits); break;
    case 2: o.x = __funnelshift_r(w[2], w[3], bits); o.y = __funnelshift_r(w[3], w[4], bits); o.z = __funnelshift_r(w[4], w[5], bits); o.w = __funnelshift_r(w[5], w[6], bits); break;
    default: o.x = __funnelshift_r(w[3], w[4], bits); o.y = __funnelshift_r(w[4], w[5], bits); o.z = __funnelshift_r(w[5], w[6], bits); o.w = __funnelshift_r(w[6], w[7], bits); break;
  }
  return o;
}

__global__ void __launch_bounds__(256) k_gather(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base,
                                                const uint4* __restrict__ recs, uint32_t n_recs) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp; r < n_recs; r += n_warps) {
    const uint4 rec = __ldg(recs + r);
    const uint32_t len = rec.w;
    if (len == 0xffffffffu) continue;
    const uint8_t* src = src_base + (((uint64_t)rec.y << 32) | rec.x);
    uint4* dst = reinterpret_cast<uint4*>(dst_base + 16ull * rec.z);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4* a = reinterpret_cast<const uint4*>(src - sh);
    const uint32_t n_dst = (len + 15u) >> 4, n_src = (sh + len + 15u) >> 4;  // aligned chunks written / read
    for (uint32_t base = 0; base < n_dst; base += 32) {
      const uint32_t c = base + lane;
      uint4 lo = make_uint4(0, 0, 0, 0), extra = make_uint4(0, 0, 0, 0);
      if (c < n_src) lo = a[c];
      if (lane == 0 && base + 32 < n_src) extra = a[base + 32];  // the chunk after this round's last: lane 31's other half
      uint4 hi;
      hi.x = __shfl_down_sync(0xffffffffu, lo.x, 1); hi.y = __shfl_down_sync(0xffffffffu, lo.y, 1);
      hi.z = __shfl_down_sync(0xffffffffu, lo.z, 1); hi.w = __shfl_down_sync(0xffffffffu, lo.w, 1);
      const uint4 e = make_uint4(__shfl_sync(0xffffffffu, extra.x, 0), __shfl_sync(0xffffffffu, extra.y, 0),
                                 __shfl_sync(0xffffffffu, extra.z, 0), __shfl_sync(0xffffffffu, extra.w, 0));
      if (lane == 31) hi = e;
      if (c < n_dst) {
        uint4 o = shift_bytes(lo, hi, sh);
        const uint32_t left = len - 16u * c;  // message bytes from this chunk on
        if (left < 16u) {  // zero the padding
          uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int keep = (int)left - 4 * k;
            w[k] &= keep >= 4 ? 0xffffffffu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
          }
          o = make_uint4(w[0], w[1], w[2], w[3]);
        }
        dst[c] = o;
      }
    }
  }
}

cudaError_t launch_gather(const uint8_t* src_base, uint8_t* dst_base, const uint4* recs, uint32_t n_recs, int sm_count,
                          cudaStream_t st) {
  if (n_recs == 0) return cudaSuccess;
  const uint32_t warps_needed = n_recs;
  uint32_t blocks = (warps_needed + 7) / 8;
  const uint32_t cap = (uint32_t)sm_count * 8;  // a resident wave: the records are few per chunk and PCIe sets the pace
  if (blocks > cap) blocks = cap;
  k_gather<<<blocks, 256, 0, st>>>(src_base, dst_base, recs, n_recs);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ key preparation (storage guest)
// The storage guest looks every storage slot up under digest_keccak(key)
// (/root/reference/circuits/risc0-storage-proof/storage-proof-circuit/storage-circuit/src/main.rs:26,
// /root/reference/trie-utils/tests/storage.rs:78).  mptv_verify_batch_hashed_keys does that on the device, inside the
// chunk pipeline and straight from the caller's key arena: one thread per proof writes the proof's (offset, length)
// key record, and for a flagged proof first hashes the key (byte loads: the arena is packed, keys are short) into
// the 32-byte slot behind the raw keys.
__global__ void __launch_bounds__(128) k_prepare_keys(const uint8_t* __restrict__ key_bytes, const uint32_t* __restrict__ key_off,
                                                      const uint32_t* __restrict__ key_len_in /* null: key_off is a prefix array */,
                                                      uint32_t key_base, const uint8_t* __restrict__ hash_key, uint64_t n_proofs,
                                                      uint8_t* __restrict__ hashed /* 32 n_proofs, 16-byte aligned */,
                                                      uint32_t hashed_off /* = hashed - key_bytes */,
                                                      uint32_t* __restrict__ off_out, uint32_t* __restrict__ len_out) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_proofs) return;
  const uint32_t o = key_off[p] - key_base, len = key_len_in ? key_len_in[p] : key_off[p + 1] - key_off[p];
  if (!hash_key[p]) { off_out[p] = o; len_out[p] = len; return; }
  const uint8_t* k = key_bytes + o;
  uint32_t lo[25], hi[25];
#pragma unroll
  for (int i = 0; i < 25; i++) { lo[i] = 0; hi[i] = 0; }
  for (uint32_t base = 0;; base += 136) {
    const uint32_t valid = len - base;  // message bytes from the start of this block
#pragma unroll
    for (int j = 0; j < 34; j++) {
      uint32_t w = 0;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const uint32_t at = 4u * j + c;
        uint32_t byte = 0;
        if (at < valid) byte = __ldg(k + base + at);
        else if (at == valid) byte = 0x01;  // pad10*1 with the Keccak delimiter (keccak.rs:7 v256)
        w |= byte << (8 * c);
      }
      if (j & 1) hi[j >> 1] ^= w; else lo[j >> 1] ^= w;
    }
    if (valid < 136u) hi[16] ^= 0x80000000u;
    keccak_f1600(lo, hi);
    if (valid < 136u) break;
  }
  uint4* out = reinterpret_cast<uint4*>(hashed + 32 * p);
  out[0] = make_uint4(lo[0], hi[0], lo[1], hi[1]);
  out[1] = make_uint4(lo[2], hi[2], lo[3], hi[3]);
  off_out[p] = hashed_off + 32u * (uint32_t)p;
  len_out[p] = 32;
}

cudaError_t launch_prepare_keys(const uint8_t* key_bytes, const uint32_t* key_off, uint32_t key_base, const uint8_t* hash_key,
                                uint64_t n_proofs, uint8_t* hashed, uint32_t hashed_off, uint32_t* off_out, uint32_t* len_out,
                                cudaStream_t st, const uint32_t* key_len_in) {
  if (n_proofs == 0) return cudaSuccess;
  k_prepare_keys<<<(unsigned)((n_proofs + 127) / 128), 128, 0, st>>>(key_bytes, key_off, key_len_in, key_base, hash_key, n_proofs, hashed,
                                                                   hashed_off, off_out, len_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ host launchers
size_t keccak_smem_bytes() { return (size_t)kStages * kKeccakThreads * kSlotBytes; }

cudaError_t kernels_init_device() {
  cudaError_t e = cudaFuncSetAttribute(k_keccak256_nodes, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)keccak_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_keccak256_leaves, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kStages * kKeccakThreads * kLeafSlotBytes);
}

// split (device word written by launch_bin_nodes, or NULL): two launches, the long nodes first with
// long_ctas CTAs per SM, then the rest at full occupancy
cudaError_t launch_keccak256_leaves(const TrieBatchDev& in, const uint4* rec, const uint32_t* node_len,
                                    const uint32_t* order, uint32_t n_nodes, uint8_t* digests, uint32_t* tile_counter,
                                    int sm_count, cudaStream_t st, const uint32_t* split, int long_ctas) {
  if (n_nodes == 0) return cudaSuccess;
  const size_t smem = (size_t)kStages * kKeccakThreads * kLeafSlotBytes;
  uint32_t n_tiles = (n_nodes + kKeccakThreads - 1) / kKeccakThreads;
  uint32_t grid = (uint32_t)sm_count * kKeccakMinBlocks;
  if (grid > n_tiles) grid = n_tiles;
  for (int part = split ? 0 : 1; part < 2; part++) {
    if (tile_counter) {
      cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(uint32_t), st);
      if (e != cudaSuccess) return e;
    }
    const uint32_t g = (split && part == 0) ? std::min<uint32_t>(grid, (uint32_t)sm_count * (uint32_t)long_ctas) : grid;
    k_keccak256_leaves<<<g, kKeccakThreads, smem, st>>>(in, rec, node_len, order, n_nodes, digests, tile_counter, split, part);
  }
  return cudaGetLastError();
}

cudaError_t launch_bin_nodes(const uint32_t* node_len, const uint32_t* ids, uint64_t n_nodes,
                             uint32_t* hist_cursor /*kBinScratchWords*/, uint32_t* order, cudaStream_t st,
                             const uint32_t* keep, unsigned long long* totals, int long_bin) {
  if (n_nodes == 0) return cudaSuccess;
  uint32_t* hist = hist_cursor;
  uint32_t* cursor = hist_cursor + kNumBins;
  cudaError_t e = cudaMemsetAsync(hist_cursor, 0, 2 * kNumBins * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  unsigned blocks = (unsigned)((n_nodes + kBinNodesPerBlock - 1) / kBinNodesPerBlock);
  if (totals) {
    e = cudaMemsetAsync(totals, 0, 16, st);
    if (e != cudaSuccess) return e;
  }
  k_bin_hist<<<blocks, 256, 0, st>>>(node_len, ids, n_nodes, hist, keep, totals);
  k_bin_scan<<<1, 32, 0, st>>>(hist, cursor, long_bin);
  k_bin_scatter<<<blocks, 256, 0, st>>>(node_len, ids, n_nodes, cursor, order, keep);
  return cudaGetLastError();
}

cudaError_t launch_keccak256_nodes(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off,
                                   const uint32_t* node_len, const uint32_t* order, uint64_t n_nodes,
                                   uint8_t* digests, uint32_t* meta, uint32_t* tile_counter, int sm_count,
                                   cudaStream_t st, const uint32_t* split, int long_ctas) {
  if (n_nodes == 0) return cudaSuccess;
  size_t smem = keccak_smem_bytes();
  uint64_t n_tiles = (n_nodes + kKeccakThreads - 1) / kKeccakThreads;
  uint64_t grid = (uint64_t)sm_count * kKeccakMinBlocks;  // persistent: one wave of resident CTAs
  if (grid > n_tiles) grid = n_tiles;
  for (int part = split ? 0 : 1; part < 2; part++) {
    if (tile_counter) {
      cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(uint32_t), st);
      if (e != cudaSuccess) return e;
    }
    const uint64_t g = (split && part == 0) ? std::min<uint64_t>(grid, (uint64_t)sm_count * long_ctas) : grid;
    k_keccak256_nodes<<<(unsigned)g, kKeccakThreads, smem, st>>>(node_bytes, byte_base, node_off, node_len, order, n_nodes,
                                                                 digests, meta, tile_counter, split, part);
  }
  return cudaGetLastError();
}

}  // namespace mptv
