// borsh_kernels.cu -- flattening borsh(MerkleProofInput) blobs ON THE DEVICE ("device flatten" mode of
// mptv_verify_borsh).  Wire format: /root/reference/crypto-ops/src/types.rs:4-9 (derive BorshSerialize: proof
// Vec<Vec<u8>>, root_hash Vec<u8>, key Vec<u8>, u32-LE length prefixes), as the prover writes it
// (/root/reference/prover/src/bin/main.rs:41,67) and the guests read it (circuits/sp1-merkle-proof/src/main.rs:5-6).
//
// When several GPUs are fed from one host, the host's memory system is what limits the host-fed path (DESIGN.md
// section 6): every byte the cores read or write counts.  With the blobs in page-locked memory the cores need not touch
// them at all: a chunk of blobs crosses PCIe as it is (one copy), and these kernels do what host_flatten.h does on
// the CPU -- walk the length prefixes (one thread per blob), lay the nodes out on 16-byte boundaries (exclusive scans
// give every blob its node indices and bytes), copy them (k_gather, device to device at HBM speed) and, after the
// verification, map each value back to its position inside the caller's blobs.
//   k_blob_count  -> per blob: node count, padded bytes, well-formed?, root_hash.len() != 32?
//   k_blob_scan   -> node_first[np + 1], byte_first[np + 1], totals and "any blob malformed" for the host
//   k_blob_emit   -> node_off / node_len / node_src, proof_first, roots, key records, gather records
//   k_blob_map    -> (status, value_off in the caller's blobs, value_len) per proof
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mptv {

namespace {

__device__ __forceinline__ uint32_t rd32(const uint8_t* p) {  // the wire format is little-endian, any alignment
  return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
}
__device__ __forceinline__ uint64_t up16d(uint64_t x) { return (x + 15) & ~15ull; }

}  // namespace

// what borsh::from_slice::<MerkleProofInput> accepts: every length fits, nothing is left over
__global__ void __launch_bounds__(128) k_blob_count(const uint8_t* __restrict__ img, const uint64_t* __restrict__ off, uint32_t np,
                                                    uint32_t* __restrict__ n_nodes, uint64_t* __restrict__ n_bytes,
                                                    uint8_t* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const uint8_t* p = img + off[i];
  const uint8_t* end = img + off[i + 1];
  bool ok = off[i + 1] >= off[i] && end - p >= 12;
  uint32_t n = 0, bad_root = 0;
  uint64_t bytes = 0;
  if (ok) {
    n = rd32(p);
    p += 4;
    ok = n <= (uint64_t)(end - p - 8) / 4;  // every node costs at least its length word
    for (uint32_t j = 0; ok && j < n; j++) {
      if (end - p < 4) { ok = false; break; }
      const uint32_t len = rd32(p);
      if ((uint64_t)(end - p - 4) < len || len > kMaxNodeLen) { ok = false; break; }
      bytes += up16d(len);
      p += 4 + len;
    }
    if (ok && end - p < 4) ok = false;
    if (ok) {
      const uint32_t rl = rd32(p);
      if ((uint64_t)(end - p - 4) < rl) ok = false;
      else { bad_root = rl != 32; p += 4 + rl; }
    }
    if (ok && end - p < 4) ok = false;
    if (ok) {
      const uint32_t kl = rd32(p);
      if ((uint64_t)(end - p - 4) < kl) ok = false;
      else { p += 4 + kl; bytes += up16d(kl); ok = p == end; }  // borsh rejects trailing bytes
    }
  }
  n_nodes[i] = ok ? n : 0;
  n_bytes[i] = ok ? bytes : 0;
  flags[i] = (uint8_t)((ok ? 1u : 0u) | (bad_root << 1));
}

// exclusive scans, one CTA (a chunk holds ~10^4 blobs); totals[0] = nodes, [1] = bytes, [2] = malformed blobs
__global__ void __launch_bounds__(1024) k_blob_scan(const uint32_t* __restrict__ n_nodes, const uint64_t* __restrict__ n_bytes,
                                                    const uint8_t* __restrict__ flags, uint32_t np, uint32_t* __restrict__ node_first,
                                                    uint64_t* __restrict__ byte_first, unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long sc[1024], sb[1024];
  __shared__ unsigned int s_bad;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const uint32_t per = (np + 1023) / 1024;
  const uint32_t s0 = min(np, tid * per), s1 = min(np, s0 + per);
  unsigned long long c = 0, b = 0;
  unsigned int bad = 0;
  for (uint32_t i = s0; i < s1; i++) { c += n_nodes[i]; b += n_bytes[i]; bad += !(flags[i] & 1); }
  sc[tid] = c; sb[tid] = b;
  if (bad) atomicAdd(&s_bad, bad);
  __syncthreads();
  if (tid == 0) {
    unsigned long long ac = 0, ab = 0;
    for (int i = 0; i < 1024; i++) {
      const unsigned long long x = sc[i], y = sb[i];
      sc[i] = ac; sb[i] = ab;
      ac += x; ab += y;
    }
    node_first[np] = (uint32_t)ac;
    byte_first[np] = ab;
    totals[0] = ac; totals[1] = ab; totals[2] = s_bad;
  }
  __syncthreads();
  c = sc[tid]; b = sb[tid];
  for (uint32_t i = s0; i < s1; i++) {
    node_first[i] = (uint32_t)c; byte_first[i] = b;
    c += n_nodes[i]; b += n_bytes[i];
  }
}

// second walk: the CSR arrays of include/mptv.h for the chunk, and one gather record per node / key
__global__ void __launch_bounds__(128) k_blob_emit(const uint8_t* __restrict__ img, const uint64_t* __restrict__ off, uint32_t np,
                                                   const uint32_t* __restrict__ node_first, const uint64_t* __restrict__ byte_first,
                                                   uint64_t arena_off /* of the byte arena inside the pack */, uint64_t blob_base,
                                                   uint64_t* __restrict__ node_off, uint32_t* __restrict__ node_len,
                                                   uint64_t* __restrict__ node_src, uint32_t* __restrict__ proof_first,
                                                   uint8_t* __restrict__ roots, uint32_t* __restrict__ key_off,
                                                   uint32_t* __restrict__ key_len, uint4* __restrict__ recs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > np) return;
  if (i == np) { proof_first[np] = node_first[np]; return; }
  const uint8_t* p = img + off[i];
  uint32_t k = node_first[i];
  const uint32_t n = node_first[i + 1] - k;
  uint64_t at = arena_off + byte_first[i];
  uint4* rec = recs + k + i;  // blob i owns records [node_first[i] + i, node_first[i + 1] + i + 1): its nodes, then its key
  proof_first[i] = k;
  p += 4;
  for (uint32_t j = 0; j < n; j++) {
    const uint32_t len = rd32(p);
    const uint64_t src = (uint64_t)(p + 4 - img);
    node_off[k] = at; node_len[k] = len; node_src[k] = blob_base + src;
    *rec++ = make_uint4((uint32_t)src, (uint32_t)(src >> 32), (uint32_t)(at >> 4), len);
    at += up16d(len);
    k++;
    p += 4 + len;
  }
  const uint32_t rl = rd32(p);
  uint8_t* r = roots + 32ull * i;
  if (rl == 32) for (int b = 0; b < 32; b++) r[b] = __ldg(p + 4 + b);
  else for (int b = 0; b < 32; b++) r[b] = 0;  // BAD_ROOT_LEN is decided from the flag; the proof is walked against zero
  p += 4 + rl;
  const uint32_t kl = rd32(p);
  const uint64_t ksrc = (uint64_t)(p + 4 - img);
  key_off[i] = (uint32_t)at; key_len[i] = kl;
  *rec = make_uint4((uint32_t)ksrc, (uint32_t)(ksrc >> 32), (uint32_t)(at >> 4), kl ? kl : 0xffffffffu);  // empty key: unused record
}

// results in place: value_off (an offset into the pack) -> position inside the caller's blobs, inside THIS proof's own
// copy of the node; the first node of the proof that holds it (the reference's lookup order); BAD_ROOT_LEN first
__global__ void __launch_bounds__(256) k_blob_map(uint32_t np, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ proof_first,
                                                  const uint64_t* __restrict__ node_off, const uint32_t* __restrict__ node_len,
                                                  const uint64_t* __restrict__ node_src, uint8_t* __restrict__ status,
                                                  uint64_t* __restrict__ value_off, uint32_t* __restrict__ value_len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  uint8_t st = status[i];
  uint64_t vo = 0;
  uint32_t vl = 0;
  if (flags[i] & 2) st = (uint8_t)kStBadRootLen;
  else if (st == kStOk) {
    vl = value_len[i];
    const uint64_t v = value_off[i];
    for (uint32_t k = proof_first[i]; k < proof_first[i + 1]; k++)
      if (node_off[k] <= v && v + vl <= node_off[k] + node_len[k]) { vo = node_src[k] + (v - node_off[k]); break; }
  }
  status[i] = st; value_off[i] = vo; value_len[i] = st == kStOk ? vl : 0u;
}

cudaError_t launch_blob_count(const uint8_t* img, const uint64_t* off, uint32_t np, uint32_t* n_nodes, uint64_t* n_bytes, uint8_t* flags,
                              uint32_t* node_first, uint64_t* byte_first, unsigned long long* totals, cudaStream_t st) {
  if (np == 0) return cudaSuccess;
  k_blob_count<<<(np + 127) / 128, 128, 0, st>>>(img, off, np, n_nodes, n_bytes, flags);
  k_blob_scan<<<1, 1024, 0, st>>>(n_nodes, n_bytes, flags, np, node_first, byte_first, totals);
  return cudaGetLastError();
}

cudaError_t launch_blob_emit(const uint8_t* img, const uint64_t* off, uint32_t np, uint32_t nn, const uint32_t* node_first,
                             const uint64_t* byte_first, uint64_t arena_off, uint64_t blob_base, uint64_t* node_off, uint32_t* node_len,
                             uint64_t* node_src, uint32_t* proof_first, uint8_t* roots, uint32_t* key_off, uint32_t* key_len,
                             uint4* recs, cudaStream_t st) {
  if (np == 0) return cudaSuccess;
  k_blob_emit<<<(np + 1 + 127) / 128, 128, 0, st>>>(img, off, np, node_first, byte_first, arena_off, blob_base, node_off, node_len,
                                                    node_src, proof_first, roots, key_off, key_len, recs);
  (void)nn;
  return cudaGetLastError();
}

cudaError_t launch_blob_map(uint32_t np, const uint8_t* flags, const uint32_t* proof_first, const uint64_t* node_off,
                            const uint32_t* node_len, const uint64_t* node_src, uint8_t* status, uint64_t* value_off,
                            uint32_t* value_len, cudaStream_t st) {
  if (np == 0) return cudaSuccess;
  k_blob_map<<<(np + 255) / 256, 256, 0, st>>>(np, flags, proof_first, node_off, node_len, node_src, status, value_off, value_len);
  return cudaGetLastError();
}

}  // namespace mptv
