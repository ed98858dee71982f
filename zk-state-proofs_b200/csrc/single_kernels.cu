// single_kernels.cu -- the latency path: a batch small enough for one CTA is verified by ONE kernel launch.
//
// Every call site of crypto_ops::verify_merkle_proof in the reference is a single-proof call
// (/root/reference/crypto-ops/src/lib.rs:8-23; /root/reference/trie-utils/tests/transaction.rs:18-22,
// tests/storage.rs:53-79 with its handful of dependent storage proofs).  Through the batch pipeline such a call costs
// one packed H2D copy, four to six kernel launches and a D2H copy -- latency, not work.  Here the host writes the
// packed batch into a page-locked mailbox that is mapped into the device's address space and launches this kernel:
//   1. the CTA copies the mailbox into shared memory (16-byte loads over PCIe, all threads at once),
//   2. one thread per node runs Keccak-256 (same keccak_f1600 as K1) and the full K2a decoder,
//   3. groups of G lanes walk the proofs (the same walk_one as K2b: the whole rule set), dependants second,
//   4. the 13-byte results and a sequence word are stored straight into the mailbox; the host polls the word.
// No copy calls, no stream synchronisation, one launch.  The rule set is verify_device.cuh, shared with the batch
// kernels, so the verdicts are the same by construction (and checked: tests/test_gpu_single.py).
#include <cuda_runtime.h>
#include <stdint.h>

#define MPTV_LDG(p) (*(p))  // the batch sits in shared memory, written earlier in this launch: plain loads
#include "keccak_f1600.cuh"
#include "kernels.h"
#include "verify_device.cuh"

namespace mptv {

// Keccak-256 of len bytes at p (shared memory, 16-byte aligned; readable up to the end of the rate block that holds
// the last byte -- the caller's layout guarantees it).  crypto-ops/src/keccak.rs:6-12: rate 136, delimiter 0x01.
__device__ void keccak256_shared(const uint8_t* p, uint32_t len, uint32_t out[8]) {
  uint32_t lo[25], hi[25];
#pragma unroll
  for (int i = 0; i < 25; i++) { lo[i] = 0; hi[i] = 0; }
  const uint32_t nb = len / 136u + 1u;
  for (uint32_t k = 0; k < nb; k++) {
    const uint2* q = reinterpret_cast<const uint2*>(p + 136u * k);
    const uint32_t valid = len - 136u * k;
    if (valid >= 136u) {
#pragma unroll
      for (int j = 0; j < 17; j++) { const uint2 w = q[j]; lo[j] ^= w.x; hi[j] ^= w.y; }
    } else {
#pragma unroll
      for (int j = 0; j < 17; j++) {
        const uint2 w = q[j];
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
    keccak_f1600(lo, hi);
  }
  out[0] = lo[0]; out[1] = hi[0]; out[2] = lo[1]; out[3] = hi[1];
  out[4] = lo[2]; out[5] = hi[2]; out[6] = lo[3]; out[7] = hi[3];
}

// the same for a short message at any alignment (a storage key): byte loads
__device__ void keccak256_bytes(const uint8_t* p, uint32_t len, uint32_t out[8]) {
  uint32_t lo[25], hi[25];
#pragma unroll
  for (int i = 0; i < 25; i++) { lo[i] = 0; hi[i] = 0; }
  for (uint32_t base = 0;; base += 136) {
    const uint32_t valid = len - base;
#pragma unroll
    for (int j = 0; j < 34; j++) {
      uint32_t w = 0;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const uint32_t at = 4u * j + c;
        uint32_t byte = 0;
        if (at < valid) byte = p[base + at];
        else if (at == valid) byte = 0x01;
        w |= byte << (8 * c);
      }
      if (j & 1) hi[j >> 1] ^= w; else lo[j >> 1] ^= w;
    }
    if (valid < 136u) hi[16] ^= 0x80000000u;
    keccak_f1600(lo, hi);
    if (valid < 136u) break;
  }
  out[0] = lo[0]; out[1] = hi[0]; out[2] = lo[1]; out[3] = hi[1];
  out[4] = lo[2]; out[5] = hi[2]; out[6] = lo[3]; out[7] = hi[3];
}

template <int G>
__global__ void __launch_bounds__(kSmallThreads, 1)
k_verify_small(const uint8_t* __restrict__ mailbox /* mapped page-locked host memory */, SmallHeader h,
               uint8_t* __restrict__ out /* mapped page-locked host memory */) {
  extern __shared__ __align__(16) uint8_t sm[];
  const uint32_t tid = threadIdx.x;
#ifdef MPTV_SMALL_TIMING
  long long tt[6];
  tt[0] = clock64();
#define MPTV_TS(i) tt[i] = clock64()
#else
#define MPTV_TS(i)
#endif
  // ---- 1. the packed batch: host memory -> shared memory, 16 bytes per thread per step
  for (uint32_t o = 16u * tid; o < h.total; o += 16u * kSmallThreads)
    *reinterpret_cast<uint4*>(sm + o) = *reinterpret_cast<const uint4*>(mailbox + o);
  // scratch behind the batch (and behind a guard of one rate block: the last node's final block is read whole)
  uint8_t* digests = sm + h.scratch;
  uint32_t* meta = reinterpret_cast<uint32_t*>(digests + 32u * h.n_nodes);
  uint64_t* value_off = reinterpret_cast<uint64_t*>(sm + h.results);
  uint32_t* value_len = reinterpret_cast<uint32_t*>(value_off + h.n_proofs);
  uint8_t* status = reinterpret_cast<uint8_t*>(value_len + h.n_proofs);
  __syncthreads();
  MPTV_TS(1);
  DeviceBatch b;
  b.node_bytes = sm + h.o_bytes; b.node_off = reinterpret_cast<const uint64_t*>(sm + h.o_off);
  b.node_len = reinterpret_cast<const uint32_t*>(sm + h.o_len); b.n_nodes = h.n_nodes;
  b.proof_first = reinterpret_cast<const uint32_t*>(sm + h.o_pf); b.n_proofs = h.n_proofs;
  b.roots = sm + h.o_roots; b.key_bytes = sm + h.o_keys; b.key_off = reinterpret_cast<const uint32_t*>(sm + h.o_koff);
  b.key_len = nullptr;
  b.root_from_proof = h.has_rfp ? reinterpret_cast<const int32_t*>(sm + h.o_rfp) : nullptr;
  b.byte_base = h.byte_base; b.node_base = h.node_base; b.key_base = h.key_base; b.proof_base = h.proof_base;
  // ---- 1b. the storage guest's digest_keccak(key) for the flagged proofs (mptv_verify_batch_hashed_keys), on the
  // threads the node loop below uses last
  if (h.has_hk) {
    uint8_t* hashed = sm + h.hk_scratch;
    uint32_t* koff2 = reinterpret_cast<uint32_t*>(hashed + 32u * h.n_proofs);
    uint32_t* klen2 = koff2 + h.n_proofs;
    const uint8_t* flags = sm + h.o_hk;
    for (uint32_t p = kSmallThreads - 1 - tid; p < h.n_proofs; p += kSmallThreads) {
      const uint32_t o = b.key_off[p] - b.key_base, len = b.key_off[p + 1] - b.key_off[p];
      if (!flags[p]) { koff2[p] = o; klen2[p] = len; continue; }
      uint32_t d[8];
      keccak256_bytes(b.key_bytes + o, len, d);
      uint4* q = reinterpret_cast<uint4*>(hashed + 32u * p);
      q[0] = make_uint4(d[0], d[1], d[2], d[3]);
      q[1] = make_uint4(d[4], d[5], d[6], d[7]);
      koff2[p] = (uint32_t)(hashed - b.key_bytes) + 32u * p;
      klen2[p] = 32;
    }
    b.key_off = koff2; b.key_len = klen2; b.key_base = 0;
  }
  // ---- 2. digest_keccak + decode_node of every supplied node, one thread each.  Consecutive nodes go to different
  // warps (schedulers): a warp instruction occupies its scheduler's 16-lane alu pipe for two cycles however few lanes
  // are active, so nodes that share a warp gain nothing and serialise wherever their code paths differ (last-block
  // padding, the decoder's branches).  One thread's Keccak chain -- 4320 alu instructions x 2 cycles = 4.4 us per
  // rate block at 1.965 GHz -- is the floor of this kernel's latency.
  for (uint32_t i = ((tid & 31u) << 2) | (tid >> 5); i < h.n_nodes; i += kSmallThreads) {  // lane l of warp w: node 4 l + w
    const uint8_t* p = b.node_bytes + (b.node_off[i] - b.byte_base);
    const uint32_t len = b.node_len[i];
    uint32_t d[8];
    keccak256_shared(p, len, d);
    uint4* o = reinterpret_cast<uint4*>(digests + 32u * i);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
    meta[i] = parse_node(p, len);
  }
  __syncthreads();
  MPTV_TS(2);
  // ---- 3. the walk: G lanes per proof, independent proofs first, then the ones whose root is an account's storage_root
  // A chain-shaped proof (what get_proof emits) is settled by ONE thread with K2f's check -- a few hundred
  // instructions instead of the cooperative walk's thousands, which matter here because every instruction of a
  // one-shot kernel is fetched cold; only what it defers goes to walk_one.
  const Group<G> g;
  __shared__ uint8_t s_need[kSmallMaxProofs];
  for (int wave = 0; wave < (h.has_rfp ? 2 : 1); wave++) {
    if (tid < h.n_proofs) {
      const bool dependent = b.root_from_proof != nullptr && b.root_from_proof[tid] >= 0;
      s_need[tid] = (dependent == (wave == 1)) && !fast_one(b, tid, dependent, digests, meta, status, value_off, value_len);
    }
    __syncthreads();
    for (uint32_t p = tid / G; p < h.n_proofs; p += kSmallThreads / G)  // uniform per group
      if (s_need[p]) walk_one<G>(b, wave, g, p, digests, meta, status, value_off, value_len);
    __syncthreads();
  }
  MPTV_TS(3);
  // ---- 4. results + sequence word -> the mailbox
  for (uint32_t i = tid; i < h.n_proofs; i += kSmallThreads) {
    reinterpret_cast<uint64_t*>(out + 16)[i] = value_off[i];
    reinterpret_cast<uint32_t*>(out + 16 + 8u * h.n_proofs)[i] = value_len[i];
    (out + 16 + 12u * h.n_proofs)[i] = status[i];
  }
  __threadfence_system();
  __syncthreads();
  MPTV_TS(4);
#ifdef MPTV_SMALL_TIMING
  if (tid == 0) for (int i = 0; i < 5; i++) reinterpret_cast<long long*>(out + 16 + 13u * kSmallMaxProofs + 16)[i] = tt[i] - tt[0];
#endif
  if (tid == 0) {
    *reinterpret_cast<volatile uint32_t*>(out) = h.seq;
    __threadfence_system();
  }
}

size_t small_smem_bytes(const SmallHeader& h) { return (size_t)h.results + 13u * h.n_proofs + 16u; }

cudaError_t small_init_device() {
  cudaError_t e = cudaFuncSetAttribute(k_verify_small<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_verify_small<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_verify_small<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallMaxSmem);
  return e;
}

cudaError_t launch_verify_small(const uint8_t* mailbox_dev, const SmallHeader& h, uint8_t* out_dev, int lanes_per_proof,
                                cudaStream_t st) {
  const size_t smem = small_smem_bytes(h);
  if (lanes_per_proof == 8) k_verify_small<8><<<1, kSmallThreads, smem, st>>>(mailbox_dev, h, out_dev);
  else if (lanes_per_proof == 16) k_verify_small<16><<<1, kSmallThreads, smem, st>>>(mailbox_dev, h, out_dev);
  else k_verify_small<32><<<1, kSmallThreads, smem, st>>>(mailbox_dev, h, out_dev);
  return cudaGetLastError();
}

}  // namespace mptv
