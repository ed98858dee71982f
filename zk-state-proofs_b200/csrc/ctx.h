// ctx.h -- internal: the context behind the opaque mptv_ctx of include/mptv.h (per-device streams,
// growable device / pinned buffers), shared by mptv_api.cu (verification) and rebuild_api.cu (tries).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/mptv.h"
#include "host_codec.h"
#include "host_flatten.h"
#include "kernels.h"

namespace mptv {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, n); want = n; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {  // page-locked staging (an async copy from / into pageable memory would stall the pipeline)
  void* p = nullptr;
  size_t cap = 0;
  unsigned flags = cudaHostAllocDefault;  // cudaHostAllocWriteCombined: written once with streaming stores, never read by the cores
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaHostAlloc(&p, want, flags);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// one pipeline slot of the host-buffer path: device copies of a chunk's inputs and outputs
constexpr size_t kPackedChunkBytes = 256 << 10;  // chunks up to this size cross PCIe as one packed copy

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;  // streamed borsh entry: the slot's chunk is through the device (blocking-sync event)
  HostBuf h_results, h_in;
  HostBuf h_wc;  // streamed borsh entry, "wc_staging": the byte regions of the chunk in write-combining memory (h_in keeps the index arrays)
  uint64_t pend_p0 = 0, pend_np = 0;  // results of this proof range are in flight into the staging buffers
  // streamed borsh entry: the chunk was flattened into h_in; these locate its index arrays there, node_src maps
  // each node back to its position in the caller's blobs (results are reported as offsets into the blobs)
  bool pend_borsh = false;
  size_t h_node_off = 0, h_node_len = 0, h_proof_first = 0;
  std::vector<uint64_t> node_src;
  std::vector<uint8_t> bad_root;  // of the chunk's blobs: root_hash.len() != 32 decides their verdict at drain time
  DevBuf node_bytes, node_off, node_len, proof_first, roots, key_bytes, key_off, rfp;
  DevBuf results, in_pack;
  DevBuf hk_flags, hk_off, hk_len;  // mptv_verify_batch_hashed_keys: the chunk's flags and the key records built on the device
  // device flatten mode of mptv_verify_borsh: the chunk's blobs as the caller wrote them, and what the walk kernels produce
  DevBuf f_img, f_off, f_nodes, f_bytes, f_flags, f_node_first, f_byte_first, f_totals;
  HostBuf f_h_off, f_h_totals;
  cudaEvent_t f_counted = nullptr;  // the chunk is on the device and counted (totals are in f_h_totals)
  uint64_t f_cs = 0, f_ce = 0;
  DevBuf digests, meta, order, bins, defer, dedup;
  void release() {
    DevBuf* all[] = {&node_bytes, &node_off, &node_len, &proof_first, &roots, &key_bytes, &key_off, &rfp,
                     &results, &in_pack, &digests, &meta, &order, &bins, &defer, &dedup, &hk_flags, &hk_off, &hk_len,
                     &f_img, &f_off, &f_nodes, &f_bytes, &f_flags, &f_node_first, &f_byte_first, &f_totals};
    for (DevBuf* b : all) b->release();
    h_results.release(); h_in.release(); h_wc.release(); f_h_off.release(); f_h_totals.release();
    if (stream) cudaStreamDestroy(stream);
    if (done) cudaEventDestroy(done);
    if (f_counted) cudaEventDestroy(f_counted);
    stream = nullptr; done = nullptr; f_counted = nullptr;
  }
};

// one of the two input stages of the host-buffer rebuild path (copy of chunk c+1 overlaps the rebuild of chunk c)
struct KvStage {
  DevBuf key_bytes, key_off, value_bytes, value_off, value_len, trie_first;
  HostBuf h_koff, h_voff, h_tfirst, h_vlen, h_keys;  // index arrays rebased to the chunk; staged small arrays
  cudaEvent_t up = nullptr;          // the stage's copies have landed
  void release() {
    DevBuf* all[] = {&key_bytes, &key_off, &value_bytes, &value_off, &value_len, &trie_first};
    for (DevBuf* b : all) b->release();
    h_koff.release(); h_voff.release(); h_tfirst.release(); h_vlen.release(); h_keys.release();
    if (up) cudaEventDestroy(up);
    up = nullptr;
  }
};

// scratch + timing of the trie rebuild path (rebuild_api.cu)
struct Rebuild {
  DevBuf rec, off, len, digests, tcount, lvl_list, sum, arena, order, bins;
  KvStage stage[2];  // host-buffer entry
  DevBuf out_roots;
  DevBuf q_trie, q_key_bytes, q_key_off, q_cnt, q_bytes, q_proof_first, q_byte_first, q_out_bytes, q_out_off, q_out_len;  // get_proof
  HostBuf h_sum, h_roots;
  cudaEvent_t ev_begin = nullptr, ev_struct = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> lvl_ev;  // 3 per level: before encode, after encode, after keccak
  uint32_t levels = 0, keccak_launches = 0, other_launches = 0;
  bool have_timing = false;
  bool leaves_in_arena = true;  // false after a fused (K1L) rebuild: hashed leaves were never written
  unsigned long long n_nodes = 0, n_hashed = 0, n_perm = 0, arena_bytes = 0;
  void release() {
    stage[0].release(); stage[1].release();
    DevBuf* all[] = {&rec, &off, &len, &digests, &tcount, &lvl_list, &sum, &arena, &order, &bins, &out_roots, &q_trie, &q_key_bytes,
                     &q_key_off, &q_cnt, &q_bytes, &q_proof_first, &q_byte_first, &q_out_bytes, &q_out_off, &q_out_len};
    for (DevBuf* b : all) b->release();
    h_sum.release(); h_roots.release();
    for (cudaEvent_t e : lvl_ev) cudaEventDestroy(e);
    lvl_ev.clear();
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_struct) cudaEventDestroy(ev_struct);
    if (ev_end) cudaEventDestroy(ev_end);
    ev_begin = ev_struct = ev_end = nullptr;
  }
};

constexpr int kSlots = 3;  // pipeline depth of the host-buffer path (H2D / kernels / D2H in flight)
// mptv_verify_borsh, host flatten: the producer may run this many chunks ahead of the device.  Three are enough when
// the pipeline has the copy engine to itself; beside the hybrid mode's 64 MB blob copies (1.2 ms each, during which a
// 14 MB staging copy waits) three slots stalled the producer for ~1 ms per device chunk.
constexpr int kBorshSlots = 6;
constexpr int kSlotsTotal = kBorshSlots + kSlots;  // the hybrid borsh mode runs a second pipeline (device flatten) on the last three

struct Device {
  int id = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;  // device-resident entry
  DevBuf digests, meta, order, bins, defer, dedup;  // scratch of the device-resident entry
  uint64_t last_unique_nodes = 0, last_unique_perm = 0;  // of the last dedup_nodes run
  Slot slot[kSlotsTotal];
  mptv_host_stats hstat2 = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // of the second pipeline's thread (hybrid mode); summed on read
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t last_stream = nullptr;
  bool have_timing = false;
  uint64_t last_nodes = 0;
  uint32_t last_keccak_launches = 0, last_other_launches = 0;
  Rebuild rb;
  // latency path: a page-locked mailbox mapped into the device's address space (inputs, then results + sequence word)
  uint8_t* mb_host = nullptr;
  uint8_t* mb_dev = nullptr;
  uint32_t mb_seq = 0;
  DedupTable dedup_tab;  // host-side candidate table of the streamed borsh entry (one chunk at a time)
  mptv_host_stats hstat = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // of the host-fed entries since the last reset
};

}  // namespace mptv

struct mptv_ctx {
  std::vector<mptv::Device> dev;
  std::string err;
  std::mutex err_mu;  // the host entries run one thread per device
  int lanes_per_proof = 0;          // 0 = auto
  uint64_t chunk_bytes = 96ull << 20;  // node bytes per pipeline chunk of the host-buffer path
  uint64_t borsh_chunk_bytes = 32ull << 20;  // borsh bytes per chunk of mptv_verify_borsh (smaller: less exposed head / tail)
  int binning = 1;
  int fused_classify = 1;  // K1 also classifies plain branches / leaves (K2a fast path)
  int dedup_nodes = 0;     // hash each DISTINCT node once (secondary mode, reported separately)
  int borsh_mode = -1;     // mptv_verify_borsh: -1 = automatic (0 on one device, 1 on several), 0 = the host flattens (and aliases),
                           // 1 = the device flattens page-locked blobs, 2 = both at once, chunks handed out from the two ends
  int hybrid_device_pct = 24;  // borsh_mode 2: share of a device's blob bytes its device-flatten pipeline may take
  int pull_pinned = 0;     // streamed borsh entry, page-locked blobs: the device gathers the placed node bytes itself (measured slower: off)
  int host_dedup = 1;      // streamed borsh entry: alias byte-identical nodes of a chunk instead of copying them again
  int wc_staging = 0;      // streamed borsh entry: stage the node bytes in write-combining page-locked memory (no cache snoops by the DMA reads)
  int latency_path = 1;    // batches that fit one CTA: one launch, mapped page-locked memory both ways (single_kernels.cu)
  int fast_walk = 1;       // K2f decides chain-shaped proofs one thread each; K2b gets the deferred rest
  int long_leaf_bin = mptv::kLongLeafBin;  // rebuild: leaves in rate-block bins >= this are hashed in their own launch ...
  int long_leaf_ctas = 1;            // ... with this many K1L CTAs (of 4 warps) per SM
  int fused_leaf_hash = 1; // rebuild: hash leaves straight from the value arena (K1L), no encode pass
};


namespace mptv {

inline int fail_cuda(mptv_ctx* c, cudaError_t e, const char* where) {
  char buf[640];
  snprintf(buf, sizeof buf, "%.500s: %s", where, cudaGetErrorString(e));
  if (c) { std::lock_guard<std::mutex> g(c->err_mu); c->err = buf; }
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();  // not sticky: the context stays usable for a smaller batch
    return MPTV_ERR_NOMEM;
  }
  return MPTV_ERR_CUDA;
}

inline int fail_msg(mptv_ctx* c, int rc, const char* msg) {
  if (c) { std::lock_guard<std::mutex> g(c->err_mu); c->err = msg; }
  return rc;
}

// A host entry that fails half way must not return while copies that read the caller's buffers (or kernels
// that will write results a later call would pick up) are still queued: wait for everything this device
// has in flight and forget the pending result blocks.
inline void quiesce(Device& d) {
  cudaSetDevice(d.id);
  if (d.stream) cudaStreamSynchronize(d.stream);
  for (Slot& s : d.slot) {  // both pipelines' slots
    if (s.stream) cudaStreamSynchronize(s.stream);
    s.pend_np = 0;
    s.pend_borsh = false;
  }
}

}  // namespace mptv

#define CK(call)                                                     \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return mptv::fail_cuda(ctx, e__, #call); \
  } while (0)
