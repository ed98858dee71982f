// dedup_kernels.cu -- optional node de-duplication in front of K1 ("dedup_nodes" option).
//
// A batch of proofs against ONE state root repeats the upper trie nodes in every proof (1 M account
// proofs against a 10 M-account trie carry the root node 1 M times).  The reference hashes every
// supplied node (crypto-ops/src/lib.rs:10-13) and so does the default pipeline -- the headline
// numbers and W_perm never deduct shared nodes.  With this option each DISTINCT node is hashed once
// and its digest / decode record is copied to its duplicates; the results are bit-identical, the
// executed Keccak-f count drops to the unique nodes', and bench.py reports such a run as a second,
// clearly labelled number.
//
// Exactness: two nodes are treated as equal only after a full byte compare (the 64-bit fingerprint
// only nominates a candidate); padding bytes beyond a node's length never take part.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mptv {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// the 16-byte chunk c of a node as four words, bytes at and beyond `len` zeroed (padding never counts)
__device__ __forceinline__ uint4 masked_chunk(const uint4* a, uint32_t c, uint32_t len) {
  uint4 x = __ldg(a + c);
  const int r = (int)len - 16 * (int)c;  // valid bytes in this chunk
  if (r < 16) {
    uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int keep = r - 4 * k;
      w[k] &= keep >= 4 ? 0xffffffffu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
    }
    x = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return x;
}

}  // namespace

// pass 1: a group of 8 lanes fingerprints the WHOLE content of a node (position-salted sum over every 16-byte
// chunk, so near-duplicates -- a tampered copy of a node -- do not collide), then one lane claims / joins
// the node's table slot and keeps the smallest node index per fingerprint.  The table is read before it
// is written: the root node of a 1 M-proof batch would otherwise queue a million atomics on one address.
__global__ void __launch_bounds__(256) k_dedup_insert(const uint8_t* __restrict__ node_bytes, uint64_t byte_base,
                                                      const uint64_t* __restrict__ node_off, const uint32_t* __restrict__ node_len,
                                                      uint32_t n_nodes, unsigned long long* __restrict__ keys,
                                                      uint32_t* __restrict__ vals, uint32_t mask, uint32_t* __restrict__ slot_of) {
  const uint32_t i = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3);  // 64-bit: 8 n_nodes threads may exceed 2^32
  const uint32_t lane = threadIdx.x & 31u, l8 = lane & 7u;
  if (i >= n_nodes) return;  // uniform per group of 8
  const uint32_t gmask = 0xffu << (lane & ~7u);
  const uint32_t len = node_len[i];
  const uint4* a = reinterpret_cast<const uint4*>(node_bytes + (node_off[i] - byte_base));
  // multilinear in the 64-bit halves with a position-dependent odd multiplier per half: a difference in any
  // single half always changes h (odd multipliers are bijections), moved chunks change it too, and the
  // two 64-bit multiplies per 16 bytes are half the work of mixing every half separately
  uint64_t h = 0;
  uint64_t m1 = (0x9e3779b97f4a7c15ull + 0xd6e8feb86659fd92ull * (uint64_t)l8) | 1ull;
  uint64_t m2 = (0xc2b2ae3d27d4eb4full + 0xa0761d6478bd642eull * (uint64_t)l8) | 1ull;
  for (uint32_t c = l8; 16 * c < len; c += 8) {
    const uint4 x = masked_chunk(a, c, len);
    const uint64_t lo = ((uint64_t)x.y << 32) | x.x, hi = ((uint64_t)x.w << 32) | x.z;
    h += lo * m1 + hi * m2;
    m1 += 8ull * 0xd6e8feb86659fd92ull;  // even steps keep the multipliers odd
    m2 += 8ull * 0xa0761d6478bd642eull;
  }
  for (int o = 4; o; o >>= 1) h ^= __shfl_xor_sync(gmask, h, o, 8);
  if (l8) return;
  uint64_t fp = mix64(h ^ ((uint64_t)len << 40));
  if (fp == 0) fp = 1;  // 0 marks an empty slot
  uint32_t s = (uint32_t)(fp >> 17) & mask;
  for (;;) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&keys[s]);
    if (cur == 0ull) cur = atomicCAS(&keys[s], 0ull, (unsigned long long)fp);
    if (cur == 0ull || cur == fp) break;
    s = (s + 1) & mask;
  }
  if (*reinterpret_cast<volatile uint32_t*>(&vals[s]) > i) atomicMin(&vals[s], i);
  slot_of[i] = s;
}

// pass 2: a group of 8 lanes per node compares it byte for byte with the candidate representative;
// dup_of[i] = representative, or i itself (unique, or a fingerprint collision)
__global__ void __launch_bounds__(256) k_dedup_resolve(const uint8_t* __restrict__ node_bytes, uint64_t byte_base,
                                                       const uint64_t* __restrict__ node_off, const uint32_t* __restrict__ node_len,
                                                       uint32_t n_nodes, const uint32_t* __restrict__ vals,
                                                       const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ dup_of) {
  const uint32_t gid = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3);  // 64-bit, as in k_dedup_insert
  const uint32_t lane = threadIdx.x & 31u, l8 = lane & 7u;
  if (gid >= n_nodes) return;  // uniform per group of 8
  const uint32_t gmask = 0xffu << (lane & ~7u);
  const uint32_t rep = vals[slot_of[gid]];
  if (rep == gid) {
    if (l8 == 0) dup_of[gid] = gid;
    return;
  }
  const uint32_t len = node_len[gid];
  bool eq = node_len[rep] == len;
  const uint4* a = reinterpret_cast<const uint4*>(node_bytes + (node_off[gid] - byte_base));
  const uint4* b = reinterpret_cast<const uint4*>(node_bytes + (node_off[rep] - byte_base));
  for (uint32_t c = l8; eq && 16 * c < len; c += 8) {  // the last chunk is masked to the node's length
    const uint4 x = masked_chunk(a, c, len), y = masked_chunk(b, c, len);
    eq = x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w;
  }
  const bool same = __all_sync(gmask, eq);
  if (l8 == 0) dup_of[gid] = same ? rep : gid;
}

// pass 3 (after K1 over the unique nodes): duplicates take their representative's digest and record
__global__ void __launch_bounds__(256) k_dedup_scatter(uint32_t n_nodes, const uint32_t* __restrict__ dup_of,
                                                       uint8_t* __restrict__ digests, uint32_t* __restrict__ meta) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const uint32_t r = dup_of[i];
  if (r == i) return;
  const uint4* s = reinterpret_cast<const uint4*>(digests + 32ull * r);
  uint4* d = reinterpret_cast<uint4*>(digests + 32ull * i);
  d[0] = s[0];
  d[1] = s[1];
  if (meta) meta[i] = meta[r];
}

cudaError_t launch_dedup_find(const uint8_t* node_bytes, uint64_t byte_base, const uint64_t* node_off, const uint32_t* node_len,
                              uint32_t n_nodes, unsigned long long* keys, uint32_t* vals, uint32_t table_size,
                              uint32_t* slot_of, uint32_t* dup_of, cudaStream_t st) {
  if (n_nodes == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(keys, 0, 8ull * table_size, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(vals, 0xff, 4ull * table_size, st);
  if (e != cudaSuccess) return e;
  const uint64_t threads = (uint64_t)n_nodes * 8;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  k_dedup_insert<<<blocks, 256, 0, st>>>(node_bytes, byte_base, node_off, node_len, n_nodes, keys, vals, table_size - 1, slot_of);
  k_dedup_resolve<<<blocks, 256, 0, st>>>(node_bytes, byte_base, node_off, node_len, n_nodes, vals, slot_of, dup_of);
  return cudaGetLastError();
}

cudaError_t launch_dedup_scatter(uint32_t n_nodes, const uint32_t* dup_of, uint8_t* digests, uint32_t* meta, cudaStream_t st) {
  if (n_nodes == 0) return cudaSuccess;
  k_dedup_scatter<<<(n_nodes + 255) / 256, 256, 0, st>>>(n_nodes, dup_of, digests, meta);
  return cudaGetLastError();
}

}  // namespace mptv
