// mptv_api.cu -- the C ABI of include/mptv.h: context, per-device memory and streams, the
// device-resident entry, and the host-buffer entry that shards a batch across the context's GPUs
// as independent proof slices (no collective, no inter-device traffic) and pipelines each slice in
// chunks.  Replaces the call boundary of crypto_ops::verify_merkle_proof
// (/root/reference/crypto-ops/src/lib.rs:8) for batches of MerkleProofInput
// (/root/reference/crypto-ops/src/types.rs:4-9).  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"
#include "host_codec.h"
#include "host_flatten.h"

using namespace mptv;

extern "C" {
static int verify_batch_impl(mptv_ctx* ctx, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out);  // defined with the C entries below
}

namespace {

int pick_lanes(const mptv_ctx* ctx, uint64_t n_nodes, uint64_t n_proofs) {
  if (ctx->lanes_per_proof == 8 || ctx->lanes_per_proof == 16 || ctx->lanes_per_proof == 32)
    return ctx->lanes_per_proof;
  double avg = n_proofs ? (double)n_nodes / (double)n_proofs : 0.0;
  return avg <= 8.5 ? 8 : (avg <= 17.0 ? 16 : 32);
}

// The whole device pipeline for one device-resident (slice of a) batch: K0 -> K1 -> K2a -> K2b.
int run_pipeline(mptv_ctx* ctx, Device& d, const DeviceBatch& b, DevBuf& digests, DevBuf& meta, DevBuf& order,
                 DevBuf& bins, DevBuf& defer, DevBuf& dedup, uint8_t* status, uint64_t* value_off, uint32_t* value_len, cudaStream_t st,
                 bool timed, mptv_host_stats* hs = nullptr /* whose launch counter (default: the device's) */) {
  CK(digests.reserve(32 * (size_t)b.n_nodes + 32));
  CK(meta.reserve(4 * (size_t)b.n_nodes + 4));
  CK(order.reserve(4 * (size_t)b.n_nodes + 4));
  CK(bins.reserve(kBinScratchWords * sizeof(uint32_t)));
  CK(defer.reserve(4 * (size_t)b.n_proofs + 8));
  uint32_t* dl = ctx->fast_walk ? defer.as<uint32_t>() : nullptr;
  if (timed) CK(cudaEventRecord(d.ev[0], st));
  const uint32_t* ord = nullptr;
  // a batch that does not even fill one wave of K1 CTAs gains nothing from binning or from dynamic tiles:
  // skip their launches (what matters for a single-proof call is latency)
  const bool one_wave = b.n_nodes <= (uint64_t)d.sm_count * kKeccakMinBlocks * kKeccakThreads / 4;
  const bool binned = ctx->binning && !one_wave;
  // optional: hash every DISTINCT node once (a secondary, separately reported mode; see dedup_kernels.cu)
  const bool dd = ctx->dedup_nodes && !one_wave && b.n_nodes <= (1ull << 30);  // the table has 2^k >= 2 n slots, k <= 31
  uint64_t n_hash = b.n_nodes;
  uint32_t* dup_of = nullptr;
  if (timed || dd) { d.last_unique_nodes = 0; d.last_unique_perm = 0; }
  if (dd) {
    uint32_t tsz = 1;
    while (tsz < 2 * b.n_nodes) tsz <<= 1;
    const size_t nn = (size_t)b.n_nodes;
    CK(dedup.reserve(16 + 12ull * tsz + 8 * nn + 64));
    uint8_t* base = dedup.as<uint8_t>();
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(base);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + 16);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + 16 + 8ull * tsz);
    uint32_t* slot_of = vals + tsz;
    dup_of = slot_of + nn;
    CK(launch_dedup_find(b.node_bytes, b.byte_base, b.node_off, b.node_len, (uint32_t)b.n_nodes, keys, vals, tsz, slot_of,
                         dup_of, st));
    // K0 bins only the representatives and counts them (per-CTA aggregated, no hot atomics)
    CK(launch_bin_nodes(b.node_len, nullptr, b.n_nodes, bins.as<uint32_t>(), order.as<uint32_t>(), st, dup_of, totals,
                        ctx->long_leaf_bin));
    unsigned long long hc[2] = {0, 0};
    CK(cudaMemcpyAsync(hc, totals, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // the number of unique nodes sizes the K1 launch
    n_hash = hc[0];
    ord = order.as<uint32_t>();
    d.last_unique_nodes = hc[0]; d.last_unique_perm = hc[1];
  } else if (binned) {
    CK(launch_bin_nodes(b.node_len, nullptr, b.n_nodes, bins.as<uint32_t>(), order.as<uint32_t>(), st, nullptr, nullptr,
                        ctx->long_leaf_bin));
    ord = order.as<uint32_t>();
  }
  // binned: nodes of more than 32 rate blocks (tx / receipt leaves of several KB) are hashed first, on their own
  const uint32_t* split = (binned || dd) ? bin_split_word(bins.as<uint32_t>()) : nullptr;
  if (timed) CK(cudaEventRecord(d.ev[1], st));
  CK(launch_keccak256_nodes(b.node_bytes, b.byte_base, b.node_off, b.node_len, ord, n_hash,
                            digests.as<uint8_t>(), ctx->fused_classify ? meta.as<uint32_t>() : nullptr,
                            one_wave ? nullptr : bins.as<uint32_t>() + 2 * kNumBins, d.sm_count, st, split, ctx->long_leaf_ctas));
  if (dd) CK(launch_dedup_scatter((uint32_t)b.n_nodes, dup_of, digests.as<uint8_t>(),
                                  ctx->fused_classify ? meta.as<uint32_t>() : nullptr, st));
  if (timed) CK(cudaEventRecord(d.ev[2], st));
  CK(launch_parse_nodes(b.node_bytes, b.byte_base, b.node_off, b.node_len, b.n_nodes, meta.as<uint32_t>(),
                        ctx->fused_classify != 0, st));
  if (timed) CK(cudaEventRecord(d.ev[3], st));
  const int G = pick_lanes(ctx, b.n_nodes, b.n_proofs);
  CK(launch_verify_walk(b, digests.as<uint8_t>(), meta.as<uint32_t>(), 0, G, status, value_off, value_len, dl,
                        d.sm_count, st));
  uint32_t other = 1 + 1 + (dl ? 1 : 0) + ((binned || dd) ? 3 : 0) + (dd ? 3 : 0);
  if (b.root_from_proof) {
    CK(launch_verify_walk(b, digests.as<uint8_t>(), meta.as<uint32_t>(), 1, G, status, value_off, value_len, dl,
                          d.sm_count, st));
    other += dl ? 2 : 1;
  }
  const uint32_t keccak_launches = b.n_nodes ? (split ? 2 : 1) : 0;
  if (timed) {
    CK(cudaEventRecord(d.ev[4], st));
    d.last_stream = st;
    d.have_timing = true;
    d.last_nodes = b.n_nodes;
    d.last_keccak_launches = keccak_launches;
    d.last_other_launches = other;
  } else {
    (hs ? *hs : d.hstat).launches += keccak_launches + other;  // the host-fed entries: kernels queued for this chunk
  }
  return MPTV_OK;
}

}  // namespace

extern "C" {

const char* mptv_strerror(int err) {
  switch (err) {
    case MPTV_OK: return "ok";
    case MPTV_ERR_ARG: return "invalid argument";
    case MPTV_ERR_CUDA: return "CUDA error (see mptv_last_error)";
    case MPTV_ERR_ALIGN: return "node not 16-byte aligned in the arena";
    case MPTV_ERR_NOMEM: return "out of memory";
    case MPTV_ERR_DEP: return "root_from_proof must reference an earlier independent proof of the same slice";
    case MPTV_ERR_NODEV: return "no usable CUDA device (there is no CPU fallback)";
  }
  return "unknown error";
}

const char* mptv_status_name(int s) {
  static const char* n[] = {"OK", "INVALID_STATE_ROOT", "ROOT_NOT_CANONICAL", "INVALID_PROOF",
                            "KEY_NOT_FOUND", "PANIC_OTHER", "BAD_ROOT_LEN", "DEPENDENCY_FAILED"};
  return (s >= 0 && s < 8) ? n[s] : "?";
}

const char* mptv_last_error(const mptv_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
int mptv_device_count(const mptv_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }

int mptv_create(const int* device_ids, int n_devices, mptv_ctx** out) {
  if (!out) return MPTV_ERR_ARG;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) return MPTV_ERR_NODEV;
  std::vector<int> ids;
  if (device_ids == nullptr || n_devices <= 0) for (int i = 0; i < visible; i++) ids.push_back(i);
  else for (int i = 0; i < n_devices; i++) {
    if (device_ids[i] < 0 || device_ids[i] >= visible) return MPTV_ERR_ARG;
    ids.push_back(device_ids[i]);
  }
  mptv_ctx* ctx = new mptv_ctx();
  ctx->dev.resize(ids.size());
  for (size_t i = 0; i < ids.size(); i++) {
    Device& d = ctx->dev[i];
    d.id = ids[i];
    cudaError_t e = cudaSetDevice(d.id);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, d.id);
    if (e == cudaSuccess && (prop.major != 10 || prop.minor != 0)) {  // the library holds sm_100a code only
      fprintf(stderr, "mptv_create: device %d is sm_%d%d, this library is built for sm_100a (B200)\n", d.id, prop.major, prop.minor);
      mptv_destroy(ctx);  // streams / events of the devices set up before this one
      return MPTV_ERR_NODEV;
    }
    if (e == cudaSuccess) { d.sm_count = prop.multiProcessorCount; e = kernels_init_device(); }
    if (e == cudaSuccess) e = trie_init_device();
    if (e == cudaSuccess) e = small_init_device();
    if (e == cudaSuccess) {
      void* hp = nullptr;
      e = cudaHostAlloc(&hp, kSmallMaxPack + kSmallOutBytes + 64, cudaHostAllocMapped | cudaHostAllocPortable);
      if (e == cudaSuccess) {
        d.mb_host = static_cast<uint8_t*>(hp);
        memset(d.mb_host, 0, kSmallMaxPack + kSmallOutBytes + 64);
        void* dp = nullptr;
        e = cudaHostGetDevicePointer(&dp, hp, 0);
        d.mb_dev = static_cast<uint8_t*>(dp);
      }
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    for (int s = 0; s < kSlotsTotal && e == cudaSuccess; s++) e = cudaStreamCreateWithFlags(&d.slot[s].stream, cudaStreamNonBlocking);
    for (int k = 0; k < 6 && e == cudaSuccess; k++) e = cudaEventCreate(&d.ev[k]);
    if (e != cudaSuccess) {
      fprintf(stderr, "mptv_create: device %d: %s\n", d.id, cudaGetErrorString(e));
      mptv_destroy(ctx);
      return MPTV_ERR_NODEV;
    }
  }
  *out = ctx;
  return MPTV_OK;
}

void mptv_destroy(mptv_ctx* ctx) {
  if (!ctx) return;
  for (Device& d : ctx->dev) {
    cudaSetDevice(d.id);
    cudaDeviceSynchronize();
    d.digests.release(); d.meta.release(); d.order.release(); d.bins.release(); d.defer.release(); d.dedup.release();
    for (int k = 0; k < kSlotsTotal; k++) d.slot[k].release();
    d.rb.release();
    if (d.mb_host) cudaFreeHost(d.mb_host);
    d.mb_host = d.mb_dev = nullptr;
    for (auto& e : d.ev) if (e) cudaEventDestroy(e);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete ctx;
}

int mptv_set_option(mptv_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return MPTV_ERR_ARG;
  if (!strcmp(name, "lanes_per_proof")) {
    if (value != 0 && value != 8 && value != 16 && value != 32) return MPTV_ERR_ARG;
    ctx->lanes_per_proof = (int)value;
  } else if (!strcmp(name, "chunk_bytes")) {
    if (value < (1 << 16)) return MPTV_ERR_ARG;
    ctx->chunk_bytes = (uint64_t)value;
  } else if (!strcmp(name, "borsh_chunk_bytes")) {
    if (value < (1 << 12) || value > (1ll << 31)) return MPTV_ERR_ARG;
    ctx->borsh_chunk_bytes = (uint64_t)value;
  } else if (!strcmp(name, "borsh_mode")) {
    if (value < -1 || value > 2) return MPTV_ERR_ARG;
    ctx->borsh_mode = (int)value;
  } else if (!strcmp(name, "hybrid_device_pct")) {
    if (value < 1 || value > 100) return MPTV_ERR_ARG;
    ctx->hybrid_device_pct = (int)value;
  } else if (!strcmp(name, "pull_pinned")) {
    ctx->pull_pinned = value ? 1 : 0;
  } else if (!strcmp(name, "wc_staging")) {
    ctx->wc_staging = value ? 1 : 0;
  } else if (!strcmp(name, "host_dedup")) {
    ctx->host_dedup = value ? 1 : 0;
  } else if (!strcmp(name, "latency_path")) {
    ctx->latency_path = value ? 1 : 0;
  } else if (!strcmp(name, "binning")) {
    ctx->binning = value ? 1 : 0;
  } else if (!strcmp(name, "fused_classify")) {
    ctx->fused_classify = value ? 1 : 0;
  } else if (!strcmp(name, "dedup_nodes")) {
    ctx->dedup_nodes = value ? 1 : 0;
  } else if (!strcmp(name, "fast_walk")) {
    ctx->fast_walk = value ? 1 : 0;
  } else if (!strcmp(name, "long_leaf_bin")) {
    if (value < 1 || value > kNumBins) return MPTV_ERR_ARG;
    ctx->long_leaf_bin = (int)value;
  } else if (!strcmp(name, "long_leaf_ctas")) {
    if (value < 1 || value > kKeccakMinBlocks) return MPTV_ERR_ARG;
    ctx->long_leaf_ctas = (int)value;
  } else if (!strcmp(name, "fused_leaf_hash")) {
    ctx->fused_leaf_hash = value ? 1 : 0;
  } else if (!strcmp(name, "l2_fetch_granularity")) {
    // device-wide hint: how many bytes an L2 miss brings in from HBM (the nodes are short, 16-byte aligned
    // runs reached in binned order, so wider fetches mostly bring in bytes of a neighbour nobody asked for yet)
    if (value != 32 && value != 64 && value != 128) return MPTV_ERR_ARG;
    for (Device& d : ctx->dev) {
      CK(cudaSetDevice(d.id));
      CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    }
  } else return MPTV_ERR_ARG;
  return MPTV_OK;
}

void* mptv_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
  return p;
}
void mptv_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ device-resident entries
int mptv_verify_batch_device(mptv_ctx* ctx, int dev_index, const mptv_batch* in, mptv_result* out, void* stream) {
  if (!ctx || !in || !out || dev_index < 0 || dev_index >= (int)ctx->dev.size()) return MPTV_ERR_ARG;
  if (in->n_nodes > 0xfffffff0ull) return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream;
  DeviceBatch b;
  b.node_bytes = in->node_bytes; b.node_off = in->node_off; b.node_len = in->node_len; b.n_nodes = in->n_nodes;
  b.proof_first = in->proof_first; b.n_proofs = in->n_proofs; b.roots = in->roots;
  b.key_bytes = in->key_bytes; b.key_off = in->key_off; b.key_len = nullptr; b.root_from_proof = in->root_from_proof;
  b.byte_base = 0; b.node_base = 0; b.key_base = 0; b.proof_base = 0;
  return run_pipeline(ctx, d, b, d.digests, d.meta, d.order, d.bins, d.defer, d.dedup, out->status, out->value_off, out->value_len,
                      st, true);
}

int mptv_keccak256_batch_device(mptv_ctx* ctx, int dev_index, const uint8_t* node_bytes, const uint64_t* node_off,
                                const uint32_t* node_len, uint64_t n_nodes, uint8_t* digests32, void* stream) {
  if (!ctx || dev_index < 0 || dev_index >= (int)ctx->dev.size()) return MPTV_ERR_ARG;
  if (n_nodes > 0xfffffff0ull) return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream;
  CK(d.order.reserve(4 * (size_t)n_nodes + 4));
  CK(d.bins.reserve(kBinScratchWords * sizeof(uint32_t)));
  CK(cudaEventRecord(d.ev[0], st));
  const uint32_t* ord = nullptr;
  if (ctx->binning) {
    CK(launch_bin_nodes(node_len, nullptr, n_nodes, d.bins.as<uint32_t>(), d.order.as<uint32_t>(), st, nullptr, nullptr,
                        ctx->long_leaf_bin));
    ord = d.order.as<uint32_t>();
  }
  CK(cudaEventRecord(d.ev[1], st));
  CK(launch_keccak256_nodes(node_bytes, 0, node_off, node_len, ord, n_nodes, digests32, nullptr,
                            d.bins.as<uint32_t>() + 2 * kNumBins, d.sm_count, st,
                            ord ? bin_split_word(d.bins.as<uint32_t>()) : nullptr, ctx->long_leaf_ctas));
  CK(cudaEventRecord(d.ev[2], st));
  CK(cudaEventRecord(d.ev[3], st));
  CK(cudaEventRecord(d.ev[4], st));
  d.last_stream = st; d.have_timing = true; d.last_nodes = n_nodes;
  d.last_keccak_launches = n_nodes ? 1 : 0; d.last_other_launches = ctx->binning ? 3 : 0;
  return MPTV_OK;
}

#ifdef MPTV_SMALL_TIMING
// diagnostic build only (tools/build_variant.py smalltiming -DMPTV_SMALL_TIMING): SM clocks at the phase boundaries of
// the last latency-path launch on device 0, relative to its start
int mptv_debug_small_clocks(mptv_ctx* ctx, long long out5[5]) {
  if (!ctx || ctx->dev.empty() || !ctx->dev[0].mb_host) return MPTV_ERR_ARG;
  memcpy(out5, ctx->dev[0].mb_host + kSmallMaxPack + 16 + 13 * kSmallMaxProofs + 16, 5 * sizeof(long long));
  return MPTV_OK;
}
#endif

int mptv_host_stats_get(mptv_ctx* ctx, mptv_host_stats* out, int reset) {
  if (!ctx || !out) return MPTV_ERR_ARG;
  memset(out, 0, sizeof *out);
  for (Device& d : ctx->dev) {
    out->chunks += d.hstat.chunks; out->nodes += d.hstat.nodes; out->nodes_aliased += d.hstat.nodes_aliased;
    out->node_bytes_supplied += d.hstat.node_bytes_supplied; out->node_bytes_placed += d.hstat.node_bytes_placed;
    out->h2d_bytes += d.hstat.h2d_bytes; out->d2h_bytes += d.hstat.d2h_bytes;
    out->launches += d.hstat.launches; out->pull_chunks += d.hstat.pull_chunks; out->device_chunks += d.hstat.device_chunks;
    out->flatten_us += d.hstat.flatten_us; out->wait_us += d.hstat.wait_us; out->map_us += d.hstat.map_us;
    out->call_us += d.hstat.call_us; out->index_us += d.hstat.index_us;
    {
      const mptv_host_stats& h2 = d.hstat2;
      out->chunks += h2.chunks; out->nodes += h2.nodes; out->node_bytes_supplied += h2.node_bytes_supplied;
      out->node_bytes_placed += h2.node_bytes_placed; out->h2d_bytes += h2.h2d_bytes; out->d2h_bytes += h2.d2h_bytes;
      out->launches += h2.launches; out->device_chunks += h2.device_chunks;
    }
    if (reset) { memset(&d.hstat, 0, sizeof d.hstat); memset(&d.hstat2, 0, sizeof d.hstat2); }
  }
  return MPTV_OK;
}

int mptv_int_issue_peak(mptv_ctx* ctx, int dev_index, int mode, double* lane_ops_per_s) {
  if (!ctx || !lane_ops_per_s || dev_index < 0 || dev_index >= (int)ctx->dev.size() || mode < 0 || mode > 8)
    return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  CK(cudaSetDevice(d.id));
  CK(d.order.reserve((size_t)d.sm_count * 8 * 256 * 4));
  CK(run_int_peak(mode, d.sm_count, d.order.as<uint32_t>(), d.stream, lane_ops_per_s));
  return MPTV_OK;
}

int mptv_last_timings(mptv_ctx* ctx, int dev_index, mptv_timings* out) {
  if (!ctx || !out || dev_index < 0 || dev_index >= (int)ctx->dev.size()) return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  memset(out, 0, sizeof *out);
  if (!d.have_timing) return MPTV_ERR_ARG;
  CK(cudaSetDevice(d.id));
  CK(cudaEventSynchronize(d.ev[4]));
  CK(cudaEventElapsedTime(&out->bin_ms, d.ev[0], d.ev[1]));
  CK(cudaEventElapsedTime(&out->keccak_ms, d.ev[1], d.ev[2]));
  CK(cudaEventElapsedTime(&out->parse_ms, d.ev[2], d.ev[3]));
  CK(cudaEventElapsedTime(&out->walk_ms, d.ev[3], d.ev[4]));
  CK(cudaEventElapsedTime(&out->total_ms, d.ev[0], d.ev[4]));
  out->n_nodes = d.last_nodes;
  out->n_unique_nodes = d.last_unique_nodes;
  out->n_unique_perm = d.last_unique_perm;
  out->keccak_launches = d.last_keccak_launches;
  out->other_launches = d.last_other_launches;
  return MPTV_OK;
}

// ------------------------------------------------------------------ host-buffer entries
}  // extern "C"

namespace {

struct Chunk { uint64_t p0, p1; };  // proof range

// end of the chunk that starts at proof s: about `chunk_bytes` node bytes (binary search over the
// monotone node offsets), never splitting a dependency group (a dependent follows its account proof)
uint64_t next_chunk_end(const mptv_batch* in, uint64_t s, uint64_t p1, uint64_t chunk_bytes) {
  const uint32_t n_total = in->proof_first[in->n_proofs];
  const uint32_t fs = in->proof_first[s];
  const uint64_t byte0 = fs < n_total ? in->node_off[fs] : 0;
  const uint64_t target = byte0 + chunk_bytes;
  uint64_t lo = s + 1, hi = p1;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) / 2;
    const uint32_t fn = in->proof_first[mid];
    const uint64_t off = fn < n_total ? in->node_off[fn] : ~0ull;
    if (off < target) lo = mid + 1; else hi = mid;
  }
  uint64_t e = lo;
  if (in->root_from_proof) while (e < p1 && in->root_from_proof[e] >= 0) e++;
  return e;
}

int validate_slice(const mptv_batch* in, uint64_t p0, uint64_t p1) {
  for (uint64_t p = p0; p < p1; p++) {
    if (in->proof_first[p + 1] < in->proof_first[p] || in->proof_first[p + 1] > in->n_nodes) return MPTV_ERR_ARG;
    if (in->key_off[p + 1] < in->key_off[p]) return MPTV_ERR_ARG;
  }
  const uint32_t n0 = in->proof_first[p0], n1 = in->proof_first[p1];
  uint64_t prev_end = 0;
  for (uint32_t i = n0; i < n1; i++) {
    const uint64_t o = in->node_off[i];
    if (o & 15) return MPTV_ERR_ALIGN;
    if (in->node_len[i] > kMaxNodeLen) return MPTV_ERR_ARG;
    // the node must lie inside the caller's arena: checked without forming o + len (a huge offset must not wrap),
    // and exactly -- only the up-to-15 padding bytes after a node, which are never hashed, may be past the end
    if (o > in->node_bytes_len || in->node_len[i] > in->node_bytes_len - o) return MPTV_ERR_ARG;
    if (o < prev_end && i > n0) return MPTV_ERR_ARG;  // nodes must be laid out in index order
    prev_end = o + in->node_len[i];
  }
  return MPTV_OK;
}

// wait for the chunk a slot is working on and hand its results to the caller's arrays
// results of a chunk travel as ONE block: [value_off u64 x np | value_len u32 x np | status u8 x np]
int drain_slot(mptv_ctx* ctx, Slot& s, mptv_result* out) {
  CK(cudaStreamSynchronize(s.stream));
  if (s.pend_np) {
    const uint8_t* h = static_cast<const uint8_t*>(s.h_results.p);
    memcpy(out->value_off + s.pend_p0, h, 8 * s.pend_np);
    memcpy(out->value_len + s.pend_p0, h + 8 * s.pend_np, 4 * s.pend_np);
    memcpy(out->status + s.pend_p0, h + 12 * s.pend_np, s.pend_np);
    s.pend_np = 0;
  }
  return MPTV_OK;
}

inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

// run one device's slice [p0, p1): chunked, multi-buffered H2D -> kernels -> D2H
int run_slice_chunks(mptv_ctx* ctx, Device& d, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out, uint64_t p0,
                     uint64_t p1) {
  if (p1 <= p0) return MPTV_OK;
  CK(cudaSetDevice(d.id));
  int rc = MPTV_OK;
  size_t ci = 0;
  for (uint64_t cs = p0; cs < p1; ci++) {
    const Chunk c = {cs, next_chunk_end(in, cs, p1, ctx->chunk_bytes)};
    cs = c.p1;
    Slot& s = d.slot[ci % kSlots];
    cudaStream_t st = s.stream;
    // argument checks of this chunk run on the host while earlier chunks are in flight
    rc = validate_slice(in, c.p0, c.p1);
    if (rc != MPTV_OK) return rc;
    rc = drain_slot(ctx, s, out);  // the slot's previous chunk must be complete before its buffers are reused
    if (rc != MPTV_OK) return rc;
    const uint64_t np = c.p1 - c.p0;
    const uint32_t n0 = in->proof_first[c.p0], n1 = in->proof_first[c.p1];
    const uint64_t nn = n1 - n0;
    const uint64_t byte0 = nn ? in->node_off[n0] : 0;
    uint64_t byte1 = nn ? in->node_off[n1 - 1] + in->node_len[n1 - 1] : 0;
    if (byte1 > in->node_bytes_len) byte1 = in->node_bytes_len;  // device buffers carry 16 spare bytes
    const uint32_t k0 = in->key_off[c.p0], k1 = in->key_off[c.p1];
    if (in->root_from_proof)
      for (uint64_t p = c.p0; p < c.p1; p++) {
        const int32_t r = in->root_from_proof[p];
        if (r >= 0 && ((uint64_t)r < c.p0 || (uint64_t)r >= p || in->root_from_proof[r] >= 0)) return MPTV_ERR_DEP;
      }
    DeviceBatch b;
    const size_t nbytes = (size_t)(byte1 - byte0), kbytes = (size_t)(k1 - k0);
    const size_t small_total = up16(nbytes + 16) + up16(8 * nn) + up16(4 * nn) + 2 * up16(4 * (np + 1)) + up16(32 * np) +
                               up16(kbytes + 16) + (in->root_from_proof ? up16(4 * np) : 0);
    if (ctx->latency_path && nn <= kSmallMaxNodes && np <= kSmallMaxProofs && small_total + 64 <= kSmallMaxPack && !ctx->dedup_nodes) {
      // latency path (single_kernels.cu): the whole chunk is packed into the mapped mailbox and verified by ONE launch
      // that reads it over PCIe and stores the results back; the host polls the sequence word -- no copy calls, no
      // stream synchronisation.  (A slot that still owes results was drained above.)
      uint8_t* h = d.mb_host;
      SmallHeader sh;
      memset(&sh, 0, sizeof sh);
      size_t o = 0;
      auto put = [&](const void* src, size_t bytes, size_t reserve_bytes) {
        if (bytes) memcpy(h + o, src, bytes);
        const uint32_t at = (uint32_t)o;
        o += up16(reserve_bytes);
        return at;
      };
      sh.o_bytes = put(in->node_bytes + byte0, nbytes, nbytes + 16);
      sh.o_off = put(in->node_off + n0, 8 * nn, 8 * nn);
      sh.o_len = put(in->node_len + n0, 4 * nn, 4 * nn);
      sh.o_pf = put(in->proof_first + c.p0, 4 * (np + 1), 4 * (np + 1));
      sh.o_roots = put(in->roots + 32 * c.p0, 32 * np, 32 * np);
      sh.o_keys = put(in->key_bytes + k0, kbytes, kbytes + 16);
      sh.o_koff = put(in->key_off + c.p0, 4 * (np + 1), 4 * (np + 1));
      sh.has_rfp = in->root_from_proof ? 1 : 0;
      if (in->root_from_proof) sh.o_rfp = put(in->root_from_proof + c.p0, 4 * np, 4 * np);
      sh.has_hk = hash_key ? 1 : 0;
      if (hash_key) sh.o_hk = put(hash_key + c.p0, np, np);
      sh.total = (uint32_t)o;
      sh.n_nodes = (uint32_t)nn; sh.n_proofs = (uint32_t)np;
      sh.scratch = (uint32_t)up16(o + 144);  // one rate block of slack: the last node's final block is read whole
      sh.results = sh.scratch + 32 * (uint32_t)nn + (uint32_t)up16(4 * nn);
      if (hash_key) { sh.hk_scratch = sh.results; sh.results += 40 * (uint32_t)np; }  // hashed keys 32 np | offsets 4 np | lengths 4 np
      sh.byte_base = byte0; sh.node_base = n0; sh.key_base = k0; sh.proof_base = c.p0;
      sh.seq = ++d.mb_seq ? d.mb_seq : ++d.mb_seq;  // never 0
      uint8_t* ho = d.mb_host + kSmallMaxPack;
      CK(launch_verify_small(d.mb_dev, sh, d.mb_dev + kSmallMaxPack, pick_lanes(ctx, nn, np), st));
      d.hstat.launches += 1; d.hstat.chunks++; d.hstat.nodes += nn; d.hstat.node_bytes_supplied += nbytes;
      d.hstat.node_bytes_placed += nbytes; d.hstat.h2d_bytes += o; d.hstat.d2h_bytes += 13 * np + 4;
      // poll the sequence word (the kernel runs ~10-20 us); fall back to the stream if it does not show up
      volatile uint32_t* seqw = reinterpret_cast<volatile uint32_t*>(ho);
      bool seen = false;
      for (uint32_t spin = 0; spin < (1u << 22); spin++) {
        if (*seqw == sh.seq) { seen = true; break; }
        _mm_pause();
      }
      if (!seen) CK(cudaStreamSynchronize(st));
      if (*seqw != sh.seq) {
        const cudaError_t e = cudaStreamSynchronize(st);
        return fail_cuda(ctx, e != cudaSuccess ? e : cudaErrorUnknown, "mptv_verify_batch: latency kernel did not complete");
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      memcpy(out->value_off + c.p0, ho + 16, 8 * np);
      memcpy(out->value_len + c.p0, ho + 16 + 8 * np, 4 * np);
      memcpy(out->status + c.p0, ho + 16 + 12 * np, np);
      continue;
    }
    if (small_total <= kPackedChunkBytes && !hash_key) {
      // small chunk (a single verify_merkle_proof call, a handful of proofs): every input array is packed
      // into one page-locked staging block and crosses PCIe as ONE copy -- latency, not bandwidth, matters
      CK(s.h_in.reserve(small_total));
      CK(s.in_pack.reserve(small_total));
      uint8_t* h = static_cast<uint8_t*>(s.h_in.p);
      uint8_t* dv = s.in_pack.as<uint8_t>();
      size_t o = 0;
      auto put = [&](const void* src, size_t bytes, size_t reserve_bytes) {
        if (bytes) memcpy(h + o, src, bytes);
        uint8_t* at = dv + o;
        o += up16(reserve_bytes);
        return at;
      };
      b.node_bytes = put(in->node_bytes + byte0, nbytes, nbytes + 16);
      b.node_off = reinterpret_cast<const uint64_t*>(put(in->node_off + n0, 8 * nn, 8 * nn));
      b.node_len = reinterpret_cast<const uint32_t*>(put(in->node_len + n0, 4 * nn, 4 * nn));
      b.proof_first = reinterpret_cast<const uint32_t*>(put(in->proof_first + c.p0, 4 * (np + 1), 4 * (np + 1)));
      b.roots = put(in->roots + 32 * c.p0, 32 * np, 32 * np);
      b.key_bytes = put(in->key_bytes + k0, kbytes, kbytes + 16);
      b.key_off = reinterpret_cast<const uint32_t*>(put(in->key_off + c.p0, 4 * (np + 1), 4 * (np + 1)));
      b.root_from_proof = in->root_from_proof
                              ? reinterpret_cast<const int32_t*>(put(in->root_from_proof + c.p0, 4 * np, 4 * np))
                              : nullptr;
      CK(cudaMemcpyAsync(dv, h, o, cudaMemcpyHostToDevice, st));
      d.hstat.h2d_bytes += o;
      b.key_len = nullptr;
      b.key_base = k0;
    } else {
      CK(s.node_bytes.reserve(nbytes + 16));
      CK(s.node_off.reserve(8 * nn + 8));
      CK(s.node_len.reserve(4 * nn + 4));
      CK(s.proof_first.reserve(4 * (np + 1)));
      CK(s.roots.reserve(32 * np));
      const size_t hashed_off = up16(kbytes + 16);  // keccak256 of the flagged keys goes behind the raw keys, same allocation
      CK(s.key_bytes.reserve(hash_key ? hashed_off + 32 * np : kbytes + 16));
      CK(s.key_off.reserve(4 * (np + 1)));
      if (in->root_from_proof) CK(s.rfp.reserve(4 * np));
      if (hash_key) {
        CK(s.hk_flags.reserve(np));
        CK(s.hk_off.reserve(4 * np));
        CK(s.hk_len.reserve(4 * np));
        CK(cudaMemcpyAsync(s.hk_flags.p, hash_key + c.p0, np, cudaMemcpyHostToDevice, st));
      }
      CK(cudaMemcpyAsync(s.node_bytes.p, in->node_bytes + byte0, nbytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(s.node_off.p, in->node_off + n0, 8 * nn, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(s.node_len.p, in->node_len + n0, 4 * nn, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(s.proof_first.p, in->proof_first + c.p0, 4 * (np + 1), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(s.roots.p, in->roots + 32 * c.p0, 32 * np, cudaMemcpyHostToDevice, st));
      if (k1 > k0) CK(cudaMemcpyAsync(s.key_bytes.p, in->key_bytes + k0, kbytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(s.key_off.p, in->key_off + c.p0, 4 * (np + 1), cudaMemcpyHostToDevice, st));
      if (in->root_from_proof)
        CK(cudaMemcpyAsync(s.rfp.p, in->root_from_proof + c.p0, 4 * np, cudaMemcpyHostToDevice, st));
      d.hstat.h2d_bytes += nbytes + 12 * nn + 8 * (np + 1) + 32 * np + kbytes + (in->root_from_proof ? 4 * np : 0);
      b.node_bytes = s.node_bytes.as<uint8_t>(); b.node_off = s.node_off.as<uint64_t>();
      b.node_len = s.node_len.as<uint32_t>();
      b.proof_first = s.proof_first.as<uint32_t>(); b.roots = s.roots.as<uint8_t>();
      b.key_bytes = s.key_bytes.as<uint8_t>(); b.key_off = s.key_off.as<uint32_t>();
      b.root_from_proof = in->root_from_proof ? s.rfp.as<int32_t>() : nullptr;
      b.key_len = nullptr;
      b.key_base = k0;
      if (hash_key) {
        // the storage guest's digest_keccak(key) on the device, straight from the caller's key arena
        CK(launch_prepare_keys(s.key_bytes.as<uint8_t>(), s.key_off.as<uint32_t>(), k0, s.hk_flags.as<uint8_t>(), np,
                               s.key_bytes.as<uint8_t>() + hashed_off, (uint32_t)hashed_off, s.hk_off.as<uint32_t>(),
                               s.hk_len.as<uint32_t>(), st));
        d.hstat.launches += 1; d.hstat.h2d_bytes += np;
        b.key_off = s.hk_off.as<uint32_t>(); b.key_len = s.hk_len.as<uint32_t>(); b.key_base = 0;
      }
    }
    b.n_nodes = nn; b.n_proofs = np;
    b.byte_base = byte0; b.node_base = n0; b.proof_base = c.p0;
    CK(s.results.reserve(13 * np + 16));
    CK(s.h_results.reserve(13 * np + 16));
    uint8_t* res = s.results.as<uint8_t>();
    rc = run_pipeline(ctx, d, b, s.digests, s.meta, s.order, s.bins, s.defer, s.dedup, res + 12 * np,
                      reinterpret_cast<uint64_t*>(res), reinterpret_cast<uint32_t*>(res + 8 * np), st, false);
    if (rc != MPTV_OK) return rc;
    CK(cudaMemcpyAsync(s.h_results.p, res, 13 * np, cudaMemcpyDeviceToHost, st));
    d.hstat.chunks++; d.hstat.nodes += nn; d.hstat.node_bytes_supplied += nbytes; d.hstat.node_bytes_placed += nbytes;
    d.hstat.d2h_bytes += 13 * np;
    s.pend_p0 = c.p0; s.pend_np = np;
  }
  for (int k = 0; k < kSlots; k++) {
    rc = drain_slot(ctx, d.slot[k], out);
    if (rc != MPTV_OK) return rc;
  }
  return MPTV_OK;
}

int run_slice(mptv_ctx* ctx, Device& d, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out, uint64_t p0, uint64_t p1) {
  const int rc = run_slice_chunks(ctx, d, in, hash_key, out, p0, p1);
  if (rc != MPTV_OK) quiesce(d);  // earlier chunks may still be reading the caller's buffers / owe results
  return rc;
}

// ------------------------------------------------------------------ streamed borsh entry
// borsh blobs -> verdicts in one pipeline: a chunk is sized by its blob bytes, walked (pass 1: shapes) and
// copied (pass 2: non-temporal stores) by a pool of host threads straight into the slot's page-locked block,
// crosses PCIe as ONE copy, runs the device pipeline, and its results come back while the host is already
// flattening the next chunk into the next slot.
struct BorshStream {
  const uint8_t* blobs;
  const uint64_t* blob_off;
  bool pinned;  // the blobs are page-locked and mapped: the devices can fetch node bytes from them directly (pull mode)
  const StorageIndex* storage = nullptr;  // set: the blobs are borsh(StorageProofInput), indexed by skim_storage_inputs
};

// Hands out the chunks of one device's blob range [front, back).  One pipeline takes them from the front; in the
// hybrid mode a second pipeline takes them from the back at the same time, so the two share the range in proportion
// to their speeds without anyone having to guess a split.
struct ChunkFeeder {
  const uint64_t* blob_off;
  uint64_t front, back;
  const uint64_t p0, p1;
  int device_pct;  // hybrid: the share of the bytes the back (device) pipeline may take
  std::mutex mu;
  ChunkFeeder(const uint64_t* off, uint64_t a, uint64_t b, int pct) : blob_off(off), front(a), back(b), p0(a), p1(b), device_pct(pct) {}
  bool take_front(uint64_t chunk_bytes, uint64_t& cs, uint64_t& ce) {
    std::lock_guard<std::mutex> g(mu);
    if (front >= back) return false;
    cs = front;
    ce = (uint64_t)(std::upper_bound(blob_off + cs + 1, blob_off + back + 1, blob_off[cs] + chunk_bytes) - blob_off);
    if (ce > cs + 1) ce--;
    if (ce > back) ce = back;
    front = ce;
    return true;
  }
  // 1 = got a chunk, 0 = nothing left, -1 = not now: the back pipeline is ahead of its share.  The device pipeline is
  // bound by PCIe and would otherwise take chunks as fast as the link moves them, starving the host pipeline's copies
  // (and every byte it takes costs 2.4 x the PCIe bytes of a byte the host pipeline flattens).
  int take_back(uint64_t chunk_bytes, uint64_t& cs, uint64_t& ce) {
    std::lock_guard<std::mutex> g(mu);
    if (front >= back) return 0;
    const uint64_t front_bytes = blob_off[front] - blob_off[p0], back_bytes = blob_off[p1] - blob_off[back];
    if (device_pct < 100 && (back_bytes + chunk_bytes / 2) * (100 - device_pct) > (front_bytes + chunk_bytes) * device_pct) return -1;
    ce = back;
    const uint64_t want = blob_off[ce] > chunk_bytes ? blob_off[ce] - chunk_bytes : 0;
    cs = (uint64_t)(std::lower_bound(blob_off + front, blob_off + ce, want) - blob_off);
    if (cs >= ce) cs = ce - 1;
    if (cs < front) cs = front;
    back = cs;
    return 1;
  }
};

// results of a finished chunk (already in the slot's page-locked result block) -> the caller's arrays, by the pool
// (on the submitter thread the same loop took ~0.3 ms a chunk and made that thread the bottleneck)
void map_results_borsh(Slot& s, mptv_result* out, WorkerPool& pool) {
  if (!s.pend_np) return;
  const uint64_t np = s.pend_np;
  const uint8_t* r = static_cast<const uint8_t*>(s.h_results.p);
  const uint64_t* voff = reinterpret_cast<const uint64_t*>(r);
  const uint32_t* vlen = reinterpret_cast<const uint32_t*>(r + 8 * np);
  const uint8_t* status = r + 12 * np;
  const uint8_t* h = static_cast<const uint8_t*>(s.h_in.p);
  const uint64_t* node_off = reinterpret_cast<const uint64_t*>(h + s.h_node_off);
  const uint32_t* node_len = reinterpret_cast<const uint32_t*>(h + s.h_node_len);
  const uint32_t* proof_first = reinterpret_cast<const uint32_t*>(h + s.h_proof_first);
  const int T = pool.size();
  const uint64_t per = (np + T - 1) / T;
  pool.run([&](int t) {
    const uint64_t lo = std::min(np, per * t), hi = std::min(np, lo + per);
    for (uint64_t i = lo; i < hi; i++) {
      const uint64_t p = s.pend_p0 + i;
      uint8_t st = status[i];
      uint64_t vo = 0;
      uint32_t vl = 0;
      if (s.bad_root[i]) st = MPTV_ST_BAD_ROOT_LEN;  // the guests' try_into().unwrap() comes first
      else if (st == MPTV_ST_OK) {
        // the value is a slice of one node of this proof (or of the identical node it aliases): report it as a
        // slice of the caller's blobs, inside THIS proof's own copy of the node -- the FIRST node of the proof whose
        // bytes hold the value (what the reference's lookup finds when a proof carries a node twice).  Shortcut: the
        // value lies strictly inside the last node -- the leaf, in any well-formed proof -- and no earlier node is
        // the same node (same offset AND length: an empty node shares the offset of whatever is placed next).
        vl = vlen[i];
        const uint32_t k0 = proof_first[i], k1 = proof_first[i + 1];
        uint32_t hit = k1;
        if (k1 > k0) {
          const uint32_t l = k1 - 1;
          if (node_off[l] < voff[i] && voff[i] + vl <= node_off[l] + node_len[l]) {
            hit = l;
            for (uint32_t j = k0; j < l; j++)
              if (node_off[j] == node_off[l] && node_len[j] == node_len[l]) { hit = j; break; }
          }
        }
        if (hit == k1)
          for (uint32_t k = k0; k < k1; k++)
            if (node_off[k] <= voff[i] && voff[i] + vl <= node_off[k] + node_len[k]) { hit = k; break; }
        if (hit < k1) vo = s.node_src[hit] + (voff[i] - node_off[hit]);
      }
      out->status[p] = st; out->value_off[p] = vo; out->value_len[p] = vl;
    }
  });
  s.pend_np = 0;
  s.pend_borsh = false;
}

// The streamed entry runs as two parties per device.  The PRODUCER (the calling thread and its worker pool) flattens
// chunk after chunk into the slots' page-locked blocks and maps finished results back to blob offsets; the SUBMITTER
// (one thread) issues each filled slot's copies and kernels and waits for the device.  The producer therefore never
// stalls on CUDA calls: the 20-odd launches and copies of a chunk overlap the flattening of the next one.
struct BorshPipe {
  enum { kFree = 0, kFilled, kInFlight, kDone };
  std::mutex mu;
  std::condition_variable cv;
  int state[kBorshSlots];
  ChunkLayout layout[kBorshSlots];
  const uint8_t* blobs_dev = nullptr;  // pull mode: the caller's blobs as the device sees them
  uint64_t produced = 0;  // chunks handed to the submitter
  bool producer_done = false;
  int err = MPTV_OK;
};

int submit_borsh_chunk(mptv_ctx* ctx, Device& d, Slot& s, const ChunkLayout& L, const uint8_t* blobs_dev) {
  cudaStream_t st = s.stream;
  const uint64_t np = L.np;
  uint8_t* h = static_cast<uint8_t*>(s.h_in.p);
  uint8_t* dv = s.in_pack.as<uint8_t>();
  // what was written: the index arrays, then each worker's used prefix of its region (neighbours merged when the
  // unused gap between them is small)
  size_t c0 = 0, c1 = L.index_end;
  uint64_t moved = 0;
  if (L.n_gather) {
    // pull mode: index arrays + gather list in one copy, then the device fetches the placed bytes itself
    CK(cudaMemcpyAsync(dv, h, L.host_total, cudaMemcpyHostToDevice, st));
    CK(launch_gather(blobs_dev, dv, reinterpret_cast<const uint4*>(dv + L.o_gather), (uint32_t)L.n_gather, d.sm_count, st));
    moved = L.host_total + L.node_bytes_placed;  // what crosses PCIe: the block, and the bytes the kernel reads
    d.hstat.launches += 1; d.hstat.pull_chunks += 1;
  } else {
    const uint8_t* src = h;  // where pack offset c0 lives on the host
    if (L.region_base) {     // the regions have a block of their own: the index arrays go first, as a copy of their own
      CK(cudaMemcpyAsync(dv, h, L.index_end, cudaMemcpyHostToDevice, st));
      moved += L.index_end;
      src = L.region_base;
      c0 = c1 = L.region_begin.empty() ? L.index_end : L.region_begin[0];
    }
    for (size_t t = 0; t <= L.region_begin.size(); t++) {
      const bool last = t == L.region_begin.size();
      if (!last && L.region_used[t] == 0) continue;
      if (!last && L.region_begin[t] <= c1 + 4096) { c1 = L.region_begin[t] + L.region_used[t]; continue; }
      if (c1 > c0) {
        CK(cudaMemcpyAsync(dv + c0, src + c0, c1 - c0, cudaMemcpyHostToDevice, st));
        moved += c1 - c0;
      }
      if (!last) { c0 = L.region_begin[t]; c1 = c0 + L.region_used[t]; }
    }
  }
  d.hstat.chunks++; d.hstat.nodes += L.nn; d.hstat.nodes_aliased += L.nodes_aliased;
  d.hstat.node_bytes_supplied += L.node_bytes_supplied; d.hstat.node_bytes_placed += L.node_bytes_placed;
  d.hstat.h2d_bytes += moved; d.hstat.d2h_bytes += 13 * np;
  DeviceBatch b;
  b.node_bytes = dv; b.node_off = reinterpret_cast<const uint64_t*>(dv + L.o_off);
  b.node_len = reinterpret_cast<const uint32_t*>(dv + L.o_len);
  b.proof_first = reinterpret_cast<const uint32_t*>(dv + L.o_pf); b.roots = dv + L.o_roots;
  b.key_bytes = dv; b.key_off = reinterpret_cast<const uint32_t*>(dv + L.o_koff);
  b.key_len = reinterpret_cast<const uint32_t*>(dv + L.o_klen);
  b.root_from_proof = nullptr;
  if (L.groups) {
    // borsh(StorageProofInput): the storage keys are the raw slots -- digest_keccak(&key) (storage-circuit/src/main.rs:26)
    // runs here, one thread per proof, and the kernels read the key records it writes; every storage proof takes its
    // root from the verified leaf of its input's account proof
    CK(launch_prepare_keys(dv, reinterpret_cast<const uint32_t*>(dv + L.o_koff), 0, dv + L.o_hk, np, dv + L.o_hashed,
                           (uint32_t)L.o_hashed, reinterpret_cast<uint32_t*>(dv + L.o_hkoff),
                           reinterpret_cast<uint32_t*>(dv + L.o_hklen), st, reinterpret_cast<const uint32_t*>(dv + L.o_klen)));
    d.hstat.launches += 1;
    b.key_off = reinterpret_cast<const uint32_t*>(dv + L.o_hkoff);
    b.key_len = reinterpret_cast<const uint32_t*>(dv + L.o_hklen);
    b.root_from_proof = reinterpret_cast<const int32_t*>(dv + L.o_rfp);
  }
  b.n_nodes = L.nn; b.n_proofs = np;
  b.byte_base = 0; b.node_base = 0; b.key_base = 0; b.proof_base = 0;
  CK(s.results.reserve(13 * np + 16));
  CK(s.h_results.reserve(13 * np + 16));
  uint8_t* res = s.results.as<uint8_t>();
  const int rc = run_pipeline(ctx, d, b, s.digests, s.meta, s.order, s.bins, s.defer, s.dedup, res + 12 * np,
                              reinterpret_cast<uint64_t*>(res), reinterpret_cast<uint32_t*>(res + 8 * np), st, false);
  if (rc != MPTV_OK) return rc;
  CK(cudaMemcpyAsync(s.h_results.p, res, 13 * np, cudaMemcpyDeviceToHost, st));
  if (!s.done) CK(cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming));
  CK(cudaEventRecord(s.done, st));
  return MPTV_OK;
}

void borsh_submitter(mptv_ctx* ctx, Device& d, BorshPipe& P) {
  int rc = MPTV_OK;
  if (cudaSetDevice(d.id) != cudaSuccess) rc = MPTV_ERR_CUDA;
  auto fail = [&](int e) {
    std::lock_guard<std::mutex> g(P.mu);
    if (P.err == MPTV_OK) P.err = e;
    P.cv.notify_all();
  };
  if (rc != MPTV_OK) { fail(rc); return; }
  uint64_t c = 0;
  for (;; c++) {
    const int k = (int)(c % kBorshSlots);
    {
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv.wait(lk, [&] { return P.err != MPTV_OK || c < P.produced || P.producer_done; });
      if (P.err != MPTV_OK) return;
      if (c >= P.produced) break;  // the producer is done and chunk c does not exist
    }
    rc = submit_borsh_chunk(ctx, d, d.slot[k], P.layout[k], P.blobs_dev);
    if (rc != MPTV_OK) { fail(rc); return; }
    {
      std::lock_guard<std::mutex> g(P.mu);
      P.state[k] = BorshPipe::kInFlight;
    }
    if (c > 0) {  // wait for the chunk before: the device always has the newest chunk queued behind it
      const int pk = (int)((c - 1) % kBorshSlots);
      if (cudaEventSynchronize(d.slot[pk].done) != cudaSuccess) { fail(fail_cuda(ctx, cudaGetLastError(), "mptv_verify_borsh")); return; }
      std::lock_guard<std::mutex> g(P.mu);
      P.state[pk] = BorshPipe::kDone;
      P.cv.notify_all();
    }
  }
  if (c > 0) {
    const int pk = (int)((c - 1) % kBorshSlots);
    if (cudaEventSynchronize(d.slot[pk].done) != cudaSuccess) { fail(fail_cuda(ctx, cudaGetLastError(), "mptv_verify_borsh")); return; }
    std::lock_guard<std::mutex> g(P.mu);
    P.state[pk] = BorshPipe::kDone;
    P.cv.notify_all();
  }
}

int run_slice_borsh_chunks(mptv_ctx* ctx, Device& d, const BorshStream& in, mptv_result* out, ChunkFeeder& feed,
                           WorkerPool& pool) {
  CK(cudaSetDevice(d.id));
  const bool alias = ctx->host_dedup != 0;
  if (alias && !d.dedup_tab.reserve((size_t)std::min<uint64_t>(ctx->borsh_chunk_bytes / 128 + 1024, 1ull << 22)))
    return fail_msg(ctx, MPTV_ERR_NOMEM, "mptv_verify_borsh: host table allocation failed");
  // pull mode: the caller's blobs are page-locked, so each chunk's staging carries only the index arrays and a gather
  // list, and a kernel fetches the placed node bytes straight from the blobs over PCIe -- the cores neither copy them
  // nor does the DMA engine read them a second time from a staging block
  const uint8_t* blobs_dev = nullptr;
  bool pull = false;
  if (in.pinned && ctx->pull_pinned && !in.storage) {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, const_cast<uint8_t*>(in.blobs), 0) == cudaSuccess && dp) { blobs_dev = static_cast<const uint8_t*>(dp); pull = true; }
    else cudaGetLastError();
  }
  BorshPipe P;
  P.blobs_dev = blobs_dev;
  for (int k = 0; k < kBorshSlots; k++) P.state[k] = BorshPipe::kFree;
  double t_wait = 0, t_map = 0, t_flat = 0, t_join = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  std::thread submitter([&] { borsh_submitter(ctx, d, P); });
  int rc = MPTV_OK;
  const char* why = nullptr;
  size_t ci = 0;
  uint64_t cs = 0, ce = 0;
  for (; rc == MPTV_OK && feed.take_front(ctx->borsh_chunk_bytes, cs, ce); ci++) {  // as many blobs as fit borsh_chunk_bytes of input
    const int k = (int)(ci % kBorshSlots);
    Slot& s = d.slot[k];
    bool have_results = false;
    double t0 = now();
    {
      // the slot's previous chunk (kBorshSlots chunks back) must be through the device before its blocks are reused
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv.wait(lk, [&] { return P.err != MPTV_OK || P.state[k] == BorshPipe::kFree || P.state[k] == BorshPipe::kDone; });
      if (P.err != MPTV_OK) { rc = P.err; break; }
      have_results = P.state[k] == BorshPipe::kDone;
    }
    double t1 = now();
    if (have_results) map_results_borsh(s, out, pool);
    double t2 = now();
    t_wait += t1 - t0; t_map += t2 - t1;
    // one pass over the chunk's blobs, straight into the slot's page-locked block; a node identical to one already
    // placed in this chunk is aliased, not copied (host_flatten.h)
    if (alias) d.dedup_tab.new_epoch();
    auto get_block = [&](size_t host_bytes, size_t dev_bytes) -> uint8_t* {
      ChunkLayout& L = P.layout[k];  // (index_end is known by now)
      const bool split = ctx->wc_staging && !pull && L.index_end > 0 && host_bytes > L.index_end;
      if (s.in_pack.reserve(dev_bytes) != cudaSuccess || s.h_in.reserve(split ? L.index_end : host_bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      if (split) {
        // the node bytes are written once with streaming stores and read only by the DMA engine: write-combining
        // memory takes the cores' caches out of those reads (no snoops); the index arrays stay cacheable, the result
        // mapping reads them back
        s.h_wc.flags = cudaHostAllocWriteCombined;
        if (s.h_wc.reserve(host_bytes - L.index_end) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        L.region_base = static_cast<uint8_t*>(s.h_wc.p) - L.index_end;
      }
      return static_cast<uint8_t*>(s.h_in.p);
    };
    if (in.storage) {
      const StorageChunkJob job = {in.blobs, in.blob_off, cs, ce, alias ? &d.dedup_tab : nullptr, in.storage};
      rc = flatten_storage_chunk(pool, job, get_block, P.layout[k], &s.node_src, &s.bad_root);
    } else {
      const BorshChunkJob job = {in.blobs, in.blob_off, cs, ce, alias ? &d.dedup_tab : nullptr, pull};
      rc = flatten_borsh_chunk(pool, job, get_block, P.layout[k], &s.node_src, &s.bad_root);
    }
    if (rc != MPTV_OK) {
      why = rc == MPTV_ERR_ARG ? "mptv_verify_borsh: a blob is not a well-formed borsh(MerkleProofInput) / borsh(StorageProofInput), or one blob needs more than 4 GiB"
                               : "mptv_verify_borsh: staging allocation failed";
      break;
    }
    t_flat += now() - t2;
    // results are indexed by proof: blob i of the MerkleProofInput stream is proof i, input i of the storage stream owns
    // proofs [proof_first[i], proof_first[i + 1])
    s.pend_p0 = in.storage ? in.storage->proof_first[cs] : cs; s.pend_np = P.layout[k].np; s.pend_borsh = true;
    s.h_node_off = P.layout[k].o_off; s.h_node_len = P.layout[k].o_len; s.h_proof_first = P.layout[k].o_pf;
    {
      std::lock_guard<std::mutex> g(P.mu);
      P.state[k] = BorshPipe::kFilled;
      P.produced++;
      P.cv.notify_all();
    }
  }
  {
    std::lock_guard<std::mutex> g(P.mu);
    if (rc != MPTV_OK && P.err == MPTV_OK) P.err = rc;  // stops the submitter
    P.producer_done = true;
    P.cv.notify_all();
  }
  double tj = now();
  submitter.join();
  t_join = now() - tj;
  // where the producer's time went (mptv_host_stats): flattening is the work, the rest is what the pipeline costs
  d.hstat.flatten_us += (uint64_t)(t_flat * 1e6); d.hstat.wait_us += (uint64_t)((t_wait + t_join) * 1e6);
  d.hstat.map_us += (uint64_t)(t_map * 1e6); d.hstat.call_us += (uint64_t)((now() - t_begin) * 1e6);
  if (rc == MPTV_OK) rc = P.err;
  if (rc != MPTV_OK) return why ? fail_msg(ctx, rc, why) : rc;
  for (int k = 0; k < kBorshSlots; k++)
    if (P.state[k] == BorshPipe::kDone) map_results_borsh(d.slot[k], out, pool);
  return MPTV_OK;
}

// ------------------------------------------------------------------ device flatten ("borsh_mode" 1)
// The blobs are page-locked: each chunk crosses PCIe as the caller wrote it and is flattened on the device
// (borsh_kernels.cu); the cores touch nothing but the chunk's offsets.  Two phases per chunk, pipelined over the slots:
//   A  copy the chunk + its offsets, walk the blobs (count), scan, read the totals back
//   B  (once A's totals are here) lay out the pack, walk again (emit), gather, verify, map the results, read them back
// A of chunk c + 1 is queued before B of chunk c, so its copy runs beside chunk c's kernels and B never waits for its
// totals.  One thread per device; no worker pool.
int borsh_device_phase_a(mptv_ctx* ctx, mptv_host_stats& hs, Slot& s, const BorshStream& in, uint64_t cs, uint64_t ce) {
  cudaStream_t st = s.stream;
  const uint64_t np = ce - cs, b0 = in.blob_off[cs], nbytes = in.blob_off[ce] - b0;
  CK(s.f_img.reserve(nbytes + 32));
  CK(s.f_off.reserve(8 * (np + 1)));
  CK(s.f_h_off.reserve(8 * (np + 1)));
  CK(s.f_nodes.reserve(4 * np + 4));
  CK(s.f_bytes.reserve(8 * np + 8));
  CK(s.f_flags.reserve(np + 16));
  CK(s.f_node_first.reserve(4 * (np + 1)));
  CK(s.f_byte_first.reserve(8 * (np + 1)));
  CK(s.f_totals.reserve(32));
  CK(s.f_h_totals.reserve(32));
  if (!s.f_counted) CK(cudaEventCreateWithFlags(&s.f_counted, cudaEventBlockingSync | cudaEventDisableTiming));
  uint64_t* ho = static_cast<uint64_t*>(s.f_h_off.p);
  for (uint64_t i = 0; i <= np; i++) ho[i] = in.blob_off[cs + i] - b0;
  CK(cudaMemcpyAsync(s.f_img.p, in.blobs + b0, nbytes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.f_off.p, ho, 8 * (np + 1), cudaMemcpyHostToDevice, st));
  CK(launch_blob_count(s.f_img.as<uint8_t>(), s.f_off.as<uint64_t>(), (uint32_t)np, s.f_nodes.as<uint32_t>(), s.f_bytes.as<uint64_t>(),
                       s.f_flags.as<uint8_t>(), s.f_node_first.as<uint32_t>(), s.f_byte_first.as<uint64_t>(),
                       s.f_totals.as<unsigned long long>(), st));
  CK(cudaMemcpyAsync(s.f_h_totals.p, s.f_totals.p, 24, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(s.f_counted, st));
  s.f_cs = cs; s.f_ce = ce;
  hs.h2d_bytes += nbytes + 8 * (np + 1); hs.launches += 2;
  return MPTV_OK;
}

int borsh_device_phase_b(mptv_ctx* ctx, Device& d, mptv_host_stats& hs, Slot& s, const BorshStream& in) {
  cudaStream_t st = s.stream;
  const uint64_t cs = s.f_cs, np = s.f_ce - s.f_cs, b0 = in.blob_off[cs];
  CK(cudaEventSynchronize(s.f_counted));
  const unsigned long long* tot = static_cast<const unsigned long long*>(s.f_h_totals.p);
  const uint64_t nn = tot[0], nbytes = tot[1];
  if (tot[2]) return fail_msg(ctx, MPTV_ERR_ARG, "mptv_verify_borsh: a blob is not a well-formed borsh(MerkleProofInput)");
  if (nn > 0xfffffff0ull) return MPTV_ERR_ARG;
  // the pack: index arrays, gather records, then the byte arena (every node on a 16-byte boundary)
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 63) & ~(size_t)63; return at; };
  const size_t o_off = take(8 * nn), o_len = take(4 * nn), o_src = take(8 * nn), o_pf = take(4 * (np + 1)), o_roots = take(32 * np),
               o_koff = take(4 * np), o_klen = take(4 * np), o_recs = take(16 * (nn + np)), o_arena = take(nbytes + 64);
  if (o > 0xffffff00ull) return fail_msg(ctx, MPTV_ERR_ARG, "mptv_verify_borsh: one blob needs more than 4 GiB");  // key offsets are 32-bit byte offsets into the pack
  CK(s.in_pack.reserve(o + 16));
  uint8_t* dv = s.in_pack.as<uint8_t>();
  CK(launch_blob_emit(s.f_img.as<uint8_t>(), s.f_off.as<uint64_t>(), (uint32_t)np, (uint32_t)nn, s.f_node_first.as<uint32_t>(),
                      s.f_byte_first.as<uint64_t>(), o_arena, b0, reinterpret_cast<uint64_t*>(dv + o_off),
                      reinterpret_cast<uint32_t*>(dv + o_len), reinterpret_cast<uint64_t*>(dv + o_src),
                      reinterpret_cast<uint32_t*>(dv + o_pf), dv + o_roots, reinterpret_cast<uint32_t*>(dv + o_koff),
                      reinterpret_cast<uint32_t*>(dv + o_klen), reinterpret_cast<uint4*>(dv + o_recs), st));
  CK(launch_gather(s.f_img.as<uint8_t>(), dv, reinterpret_cast<const uint4*>(dv + o_recs), (uint32_t)(nn + np), d.sm_count, st));
  DeviceBatch b;
  b.node_bytes = dv; b.node_off = reinterpret_cast<const uint64_t*>(dv + o_off);
  b.node_len = reinterpret_cast<const uint32_t*>(dv + o_len);
  b.proof_first = reinterpret_cast<const uint32_t*>(dv + o_pf); b.roots = dv + o_roots;
  b.key_bytes = dv; b.key_off = reinterpret_cast<const uint32_t*>(dv + o_koff);
  b.key_len = reinterpret_cast<const uint32_t*>(dv + o_klen);
  b.root_from_proof = nullptr;
  b.n_nodes = nn; b.n_proofs = np;
  b.byte_base = 0; b.node_base = 0; b.key_base = 0; b.proof_base = 0;
  CK(s.results.reserve(13 * np + 16));
  CK(s.h_results.reserve(13 * np + 16));
  uint8_t* res = s.results.as<uint8_t>();
  const int rc = run_pipeline(ctx, d, b, s.digests, s.meta, s.order, s.bins, s.defer, s.dedup, res + 12 * np,
                              reinterpret_cast<uint64_t*>(res), reinterpret_cast<uint32_t*>(res + 8 * np), st, false, &hs);
  if (rc != MPTV_OK) return rc;
  CK(launch_blob_map((uint32_t)np, s.f_flags.as<uint8_t>(), b.proof_first, b.node_off, b.node_len,
                     reinterpret_cast<const uint64_t*>(dv + o_src), res + 12 * np, reinterpret_cast<uint64_t*>(res),
                     reinterpret_cast<uint32_t*>(res + 8 * np), st));
  CK(cudaMemcpyAsync(s.h_results.p, res, 13 * np, cudaMemcpyDeviceToHost, st));
  if (!s.done) CK(cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming));
  CK(cudaEventRecord(s.done, st));
  s.pend_p0 = cs; s.pend_np = np;
  hs.chunks++; hs.device_chunks++; hs.nodes += nn; hs.node_bytes_supplied += nbytes;
  hs.node_bytes_placed += nbytes; hs.d2h_bytes += 13 * np + 24; hs.launches += 3;
  return MPTV_OK;
}

int borsh_device_drain(mptv_ctx* ctx, Slot& s, mptv_result* out) {
  if (!s.pend_np) return MPTV_OK;
  CK(cudaEventSynchronize(s.done));
  const uint8_t* h = static_cast<const uint8_t*>(s.h_results.p);
  memcpy(out->value_off + s.pend_p0, h, 8 * s.pend_np);
  memcpy(out->value_len + s.pend_p0, h + 8 * s.pend_np, 4 * s.pend_np);
  memcpy(out->status + s.pend_p0, h + 12 * s.pend_np, s.pend_np);
  s.pend_np = 0;
  return MPTV_OK;
}

int run_slice_borsh_device(mptv_ctx* ctx, Device& d, const BorshStream& in, mptv_result* out, ChunkFeeder& feed, bool second) {
  CK(cudaSetDevice(d.id));
  const int slot0 = second ? kBorshSlots : 0;  // the hybrid mode's second pipeline: its own slots, chunks from the back
  mptv_host_stats& hs = second ? d.hstat2 : d.hstat;
  // Chunks of twice borsh_chunk_bytes: there is no host stage whose head a small chunk would shorten, and this thread's
  // ~25 CUDA calls and two event waits per chunk (~0.3 ms) must stay well below the chunk's copy time or the copy
  // engine runs dry between chunks (measured, 1 M config-2 proofs: 16 MB 91 ms, 32 MB 65 ms, 64 MB 62 ms).
  const uint64_t chunk_bytes = 2 * ctx->borsh_chunk_bytes;
  const double t_begin = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  Slot* prev = nullptr;
  size_t ci = 0;
  int rc = MPTV_OK;
  uint64_t cs = 0, ce = 0;
  for (; rc == MPTV_OK; ci++) {
    if (second) {
      int got;
      while ((got = feed.take_back(chunk_bytes, cs, ce)) < 0) std::this_thread::sleep_for(std::chrono::microseconds(100));
      if (!got) break;
    } else if (!feed.take_front(chunk_bytes, cs, ce)) break;
    Slot& s = d.slot[slot0 + ci % kSlots];
    rc = borsh_device_drain(ctx, s, out);  // the slot's chunk of three chunks ago
    if (rc == MPTV_OK) rc = borsh_device_phase_a(ctx, hs, s, in, cs, ce);
    if (rc == MPTV_OK && prev) rc = borsh_device_phase_b(ctx, d, hs, *prev, in);
    prev = &s;
  }
  if (rc == MPTV_OK && prev) rc = borsh_device_phase_b(ctx, d, hs, *prev, in);
  for (int k = 0; k < kSlots && rc == MPTV_OK; k++) rc = borsh_device_drain(ctx, d.slot[slot0 + k], out);
  if (!second) hs.call_us += (uint64_t)((std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_begin) * 1e6);
  return rc;
}

int run_slice_borsh(mptv_ctx* ctx, Device& d, const BorshStream& in, mptv_result* out, uint64_t p0, uint64_t p1, int n_threads) {
  if (p1 <= p0) return MPTV_OK;
  ChunkFeeder feed(in.blob_off, p0, p1, ctx->hybrid_device_pct);
  int rc = MPTV_OK;
  // automatic: one device has the host's cores and memory to itself and the host flatten moves the fewest PCIe bytes;
  // several devices of one context share them, and the device flatten costs the host one DMA read per byte
  const int mode = in.storage ? 0 : ctx->borsh_mode >= 0 ? ctx->borsh_mode : (ctx->dev.size() > 1 ? 1 : 0);  // (the storage stream has the host flatten only)
  if (mode == 1 && in.pinned) {
    rc = run_slice_borsh_device(ctx, d, in, out, feed, false);
  } else if (mode == 2 && in.pinned) {
    // hybrid: the host flattens (and aliases) chunks from the front of the range while the device flattens chunks from
    // its back -- the first is bound by the cores and the host's memory, the second by PCIe, so together they use both
    int rc_b = MPTV_OK;
    std::thread second([&] { rc_b = run_slice_borsh_device(ctx, d, in, out, feed, true); });
    {
      WorkerPool pool(n_threads > 1 ? n_threads - 1 : 1);  // one core less: the second pipeline's thread issues CUDA calls too
      rc = run_slice_borsh_chunks(ctx, d, in, out, feed, pool);
    }
    second.join();
    if (rc == MPTV_OK) rc = rc_b;
  } else {
    WorkerPool pool(n_threads);
    rc = run_slice_borsh_chunks(ctx, d, in, out, feed, pool);
  }
  if (rc != MPTV_OK) quiesce(d);
  return rc;
}

// A handful of inputs (one verify_merkle_proof call's MerkleProofInput, one StorageProofInput): no worker pool, no
// submitter thread, no chunk pipeline -- the blobs are flattened on the calling thread into a small block and handed to
// the batch entry, which verifies them with ONE kernel launch through the mapped mailbox (the latency path).
// storage: the index of the inputs when the blobs are borsh(StorageProofInput), else null.
constexpr uint64_t kSmallBorshBytes = 40 << 10;
int run_small_borsh(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, const StorageIndex* storage,
                    mptv_result* out) {
  WorkerPool pool(1);
  ChunkLayout L;
  std::vector<uint64_t> node_src;
  std::vector<uint8_t> bad;
  uint8_t* block = nullptr;
  size_t block_cap = 0;
  auto get_block = [&](size_t host_bytes, size_t) -> uint8_t* {
    block_cap = (host_bytes + 127) & ~(size_t)63;
    block = static_cast<uint8_t*>(aligned_alloc(64, block_cap));
    return block;
  };
  struct Free { uint8_t*& p; ~Free() { free(p); } } guard{block};
  int rc;
  if (storage) {
    const StorageChunkJob job = {blobs, blob_off, 0, n, nullptr, storage};
    rc = flatten_storage_chunk(pool, job, get_block, L, &node_src, &bad);
  } else {
    const BorshChunkJob job = {blobs, blob_off, 0, n, nullptr, false};
    rc = flatten_borsh_chunk(pool, job, get_block, L, &node_src, &bad);
  }
  if (rc != MPTV_OK)
    return fail_msg(ctx, rc, rc == MPTV_ERR_ARG ? "mptv_verify_borsh: a blob is not well-formed borsh" : "mptv_verify_borsh: allocation failed");
  const uint64_t np = L.np;
  const uint64_t* node_off = reinterpret_cast<const uint64_t*>(block + L.o_off);
  const uint32_t* node_len = reinterpret_cast<const uint32_t*>(block + L.o_len);
  const uint32_t* proof_first = reinterpret_cast<const uint32_t*>(block + L.o_pf);
  const uint32_t* koff = reinterpret_cast<const uint32_t*>(block + L.o_koff);
  const uint32_t* klen = reinterpret_cast<const uint32_t*>(block + L.o_klen);
  // the public batch form keeps the keys as a prefix array
  std::vector<uint8_t> kb;
  std::vector<uint32_t> ko(np + 1, 0);
  for (uint64_t p = 0; p < np; p++) {
    kb.insert(kb.end(), block + koff[p], block + koff[p] + klen[p]);
    ko[p + 1] = (uint32_t)kb.size();
  }
  kb.resize(kb.size() + 16, 0);
  mptv_batch b;
  memset(&b, 0, sizeof b);
  b.node_bytes = block; b.node_bytes_len = block_cap;
  b.node_off = node_off; b.node_len = node_len; b.n_nodes = L.nn;
  b.proof_first = proof_first; b.n_proofs = np; b.roots = block + L.o_roots;
  b.key_bytes = kb.data(); b.key_off = ko.data();
  b.root_from_proof = L.groups ? reinterpret_cast<const int32_t*>(block + L.o_rfp) : nullptr;
  std::vector<uint8_t> st(np);
  std::vector<uint64_t> vo(np);
  std::vector<uint32_t> vl(np);
  mptv_result r = {st.data(), vo.data(), vl.data()};
  rc = verify_batch_impl(ctx, &b, L.groups ? block + L.o_hk : nullptr, &r);
  if (rc != MPTV_OK) return rc;
  for (uint64_t p = 0; p < np; p++) {
    uint8_t s8 = st[p];
    uint64_t v = 0;
    uint32_t l = 0;
    if (bad[p]) s8 = MPTV_ST_BAD_ROOT_LEN;  // the guests' try_into().unwrap() comes first
    else if (s8 == MPTV_ST_OK) {
      l = vl[p];
      // the first node of the proof that holds the value, reported inside this proof's own blob
      for (uint32_t k = proof_first[p]; k < proof_first[p + 1]; k++)
        if (node_off[k] <= vo[p] && vo[p] + l <= node_off[k] + node_len[k]) { v = node_src[k] + (vo[p] - node_off[k]); break; }
    }
    out->status[p] = s8; out->value_off[p] = v; out->value_len[p] = l;
  }
  return MPTV_OK;
}

int verify_borsh_run(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads, mptv_result* out) {
  if (!ctx || !out) return MPTV_ERR_ARG;
  if (n == 0) return MPTV_OK;
  if (!blobs || !blob_off || !out->status || !out->value_off || !out->value_len) return MPTV_ERR_ARG;
  for (uint64_t i = 0; i < n; i++)
    if (blob_off[i + 1] < blob_off[i]) return MPTV_ERR_ARG;  // the chunking below searches the offsets
  if (ctx->latency_path && n <= kSmallMaxProofs && blob_off[n] - blob_off[0] <= kSmallBorshBytes)
    return run_small_borsh(ctx, blobs, blob_off, n, nullptr, out);
  // all cores but two per device: the submitter thread of each device needs a share of a core for its CUDA calls, and
  // the pool's barriers cost more than a core's worth of work as soon as one of its threads is descheduled (16-core
  // box, 1 M config-2 proofs: 48.9 ms with 15 threads, 46.0 ms with 14)
  if (n_threads <= 0) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nd = (unsigned)ctx->dev.size();
    n_threads = (int)std::max(1u, std::min(32u * nd, hw >= 8 * nd ? hw - 2 * nd : (hw > 2 * nd ? hw - nd : hw)));
  }
  // page-locked input?  (mptv_alloc_pinned / cudaHostAlloc / cudaHostRegister: first and last byte must both be)
  bool pinned = false;
  {
    cudaPointerAttributes a0, a1;
    const uint64_t last = blob_off[n] > blob_off[0] ? blob_off[n] - 1 : blob_off[0];
    if (cudaPointerGetAttributes(&a0, blobs + blob_off[0]) == cudaSuccess && cudaPointerGetAttributes(&a1, blobs + last) == cudaSuccess)
      pinned = a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  const BorshStream in = {blobs, blob_off, pinned};
  const int nd = (int)ctx->dev.size();
  std::vector<uint64_t> cut(nd + 1, 0);
  cut[nd] = n;
  const uint64_t total = blob_off[n] - blob_off[0];
  for (int k = 1; k < nd; k++)
    cut[k] = (uint64_t)(std::lower_bound(blob_off, blob_off + n + 1, blob_off[0] + total / nd * k) - blob_off);
  for (int k = 1; k <= nd; k++) if (cut[k] < cut[k - 1]) cut[k] = cut[k - 1];
  std::vector<int> rcs(nd, MPTV_OK);
  if (nd == 1) {
    rcs[0] = run_slice_borsh(ctx, ctx->dev[0], in, out, cut[0], cut[1], n_threads);
  } else {
    std::vector<std::thread> th;
    const int per = std::max(1, n_threads / nd);
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] { rcs[k] = run_slice_borsh(ctx, ctx->dev[k], in, out, cut[k], cut[k + 1], per); });
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nd; k++) if (rcs[k] != MPTV_OK) return rcs[k];
  return MPTV_OK;
}

// ------------------------------------------------------------------ borsh(StorageProofInput) stream
int verify_storage_borsh_run(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads,
                             uint64_t* proof_first, uint8_t* input_status, uint64_t results_cap, mptv_result* out) {
  if (!ctx || !proof_first) return MPTV_ERR_ARG;
  proof_first[0] = 0;
  if (n == 0) return MPTV_OK;
  if (!blobs || !blob_off) return MPTV_ERR_ARG;
  for (uint64_t i = 0; i < n; i++)
    if (blob_off[i + 1] < blob_off[i]) return MPTV_ERR_ARG;
  const int nd = (int)ctx->dev.size();
  if (n_threads <= 0) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::max(1u, std::min(32u * nd, hw >= 8u * nd ? hw - 2 * nd : (hw > 2u * nd ? hw - nd : hw)));
  }
  // pass 1: the length prefixes of every input -> how many proofs (and nodes) the guest verifies for it.  (Indexing
  // chunk by chunk inside the stream, so that the flattener finds the prefixes in cache, was measured and is no
  // faster: 247 vs 248 ms for the 1 M inputs of config 3 -- the flattener's own prefetching already hides those misses,
  // and 341 small index passes cost more than one large one.)
  StorageIndex idx;
  {
    const auto t0 = std::chrono::steady_clock::now();
    WorkerPool pool(n < 64 ? 1 : n_threads);
    const int rc = skim_storage_inputs(pool, blobs, blob_off, n, idx);
    if (rc != MPTV_OK) return fail_msg(ctx, rc, "mptv_verify_storage_borsh: a blob is not a well-formed borsh(StorageProofInput)");
    ctx->dev[0].hstat.index_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
  }
  memcpy(proof_first, idx.proof_first.data(), 8 * (n + 1));
  if (results_cap < idx.proof_first[n] || !out || !out->status || !out->value_off || !out->value_len) return MPTV_ERR_NOMEM;  // proof_first[n] = what is required
  if (ctx->latency_path && idx.proof_first[n] <= kSmallMaxProofs && blob_off[n] - blob_off[0] <= kSmallBorshBytes) {
    const int rc = run_small_borsh(ctx, blobs, blob_off, n, &idx, out);  // one StorageProofInput: one kernel launch
    if (rc != MPTV_OK) return rc;
  } else {
  // pass 2: the stream, cut over the devices by blob bytes at input boundaries
  const BorshStream in = {blobs, blob_off, false, &idx};
  std::vector<uint64_t> cut(nd + 1, 0);
  cut[nd] = n;
  const uint64_t total = blob_off[n] - blob_off[0];
  for (int k = 1; k < nd; k++)
    cut[k] = (uint64_t)(std::lower_bound(blob_off, blob_off + n + 1, blob_off[0] + total / nd * k) - blob_off);
  for (int k = 1; k <= nd; k++) if (cut[k] < cut[k - 1]) cut[k] = cut[k - 1];
  std::vector<int> rcs(nd, MPTV_OK);
  if (nd == 1) {
    rcs[0] = run_slice_borsh(ctx, ctx->dev[0], in, out, cut[0], cut[1], n_threads);
  } else {
    std::vector<std::thread> th;
    const int per = std::max(1, n_threads / nd);
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] { rcs[k] = run_slice_borsh(ctx, ctx->dev[k], in, out, cut[k], cut[k + 1], per); });
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nd; k++) if (rcs[k] != MPTV_OK) return rcs[k];
  }
  // the guest's outcome per input: the first proof that fails, in its order; an account leaf that is not an Account
  // (decode_exact(...).unwrap(), main.rs:15) fails the input even when it carries no storage proof
  if (input_status) {
    WorkerPool pool(n < 4096 ? 1 : n_threads);
    const int T = pool.size();
    const uint64_t per = (n + T - 1) / T;
    pool.run([&](int t) {
      const uint64_t lo = std::min(n, per * t), hi = std::min(n, lo + per);
      for (uint64_t i = lo; i < hi; i++) {
        const uint64_t a = idx.proof_first[i], e = idx.proof_first[i + 1];
        uint8_t st = out->status[a];
        // (with storage proofs behind it the device has made the same check: they carry MPTV_ST_DEP_FAILED)
        if (st == MPTV_ST_OK && e == a + 1 && !mptv_account_storage_root(blobs + out->value_off[a], out->value_len[a], nullptr))
          st = MPTV_ST_DEP_FAILED;
        for (uint64_t q = a + 1; st == MPTV_ST_OK && q < e; q++) st = out->status[q];
        input_status[i] = st;
      }
    });
  }
  return MPTV_OK;
}

}  // namespace

extern "C" {

int mptv_verify_storage_borsh(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n_inputs, int n_threads,
                              uint64_t* proof_first, uint8_t* input_status, uint64_t results_cap, mptv_result* out) {
  try {
    return verify_storage_borsh_run(ctx, blobs, blob_off, n_inputs, n_threads, proof_first, input_status, results_cap, out);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

int mptv_verify_borsh(mptv_ctx* ctx, const uint8_t* blobs, const uint64_t* blob_off, uint64_t n, int n_threads,
                      mptv_result* out) {
  try {  // host tables sized by n; the C ABI never throws
    return verify_borsh_run(ctx, blobs, blob_off, n, n_threads, out);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

static int verify_batch_impl(mptv_ctx* ctx, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out) {
  if (!ctx || !in || !out) return MPTV_ERR_ARG;
  if (in->n_proofs == 0) return MPTV_OK;
  if (!in->node_off || !in->node_len || !in->proof_first || !in->roots || !in->key_off || !out->status ||
      !out->value_off || !out->value_len)
    return MPTV_ERR_ARG;
  if (in->n_nodes > 0xfffffff0ull || in->proof_first[in->n_proofs] > in->n_nodes) return MPTV_ERR_ARG;
  const int nd = (int)ctx->dev.size();
  // slice boundaries: equal shares of the Keccak-f count (~ node bytes), never inside a dependency group
  std::vector<uint64_t> cut(nd + 1, 0);
  cut[nd] = in->n_proofs;
  if (nd > 1) {
    const uint32_t n_total = in->proof_first[in->n_proofs];
    const uint64_t total = n_total ? in->node_off[n_total - 1] + in->node_len[n_total - 1] : 0;
    uint64_t p = 0;
    for (int k = 1; k < nd; k++) {
      const uint64_t target = total / nd * k;
      // binary search the first proof whose first node starts at or after `target`
      uint64_t lo = p, hi = in->n_proofs;
      while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        const uint32_t fn = in->proof_first[mid];
        const uint64_t off = fn < n_total ? in->node_off[fn] : total;
        if (off < target) lo = mid + 1; else hi = mid;
      }
      p = lo;
      if (in->root_from_proof) while (p < in->n_proofs && in->root_from_proof[p] >= 0) p++;
      cut[k] = p;
    }
  }
  std::vector<int> rcs(nd, MPTV_OK);
  if (nd == 1) {
    rcs[0] = run_slice(ctx, ctx->dev[0], in, hash_key, out, cut[0], cut[1]);
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] { rcs[k] = run_slice(ctx, ctx->dev[k], in, hash_key, out, cut[k], cut[k + 1]); });
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nd; k++) if (rcs[k] != MPTV_OK) return rcs[k];
  return MPTV_OK;
}

int mptv_verify_batch(mptv_ctx* ctx, const mptv_batch* in, mptv_result* out) {
  try {  // per-device threads / slice tables; the C ABI never throws
    return verify_batch_impl(ctx, in, nullptr, out);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// The flags travel with each chunk and the flagged keys are hashed on the device inside the chunk pipeline
// (k_prepare_keys), straight from the caller's key arena: no host-side key table, no extra pass over the batch.
int mptv_verify_batch_hashed_keys(mptv_ctx* ctx, const mptv_batch* in, const uint8_t* hash_key, mptv_result* out) {
  try {
    return verify_batch_impl(ctx, in, hash_key, out);
  } catch (...) {
    return MPTV_ERR_NOMEM;
  }
}

// one device's share [i0, i1) of a host arena: the byte range its nodes span goes up as one copy
static int keccak256_slice(mptv_ctx* ctx, Device& d, const uint8_t* node_bytes, uint64_t node_bytes_len, const uint64_t* node_off,
                           const uint32_t* node_len, uint64_t i0, uint64_t i1, uint8_t* digests32) {
  if (i1 <= i0) return MPTV_OK;
  const uint64_t n = i1 - i0;
  uint64_t lo = ~0ull, hi = 0;
  for (uint64_t i = i0; i < i1; i++) {
    lo = std::min(lo, node_off[i]);
    hi = std::max(hi, node_off[i] + node_len[i]);
  }
  Slot& s = d.slot[0];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = s.stream;
  const uint64_t span = ((hi - lo) + 15) & ~15ull;
  const uint64_t avail = std::min(span, node_bytes_len - lo);  // the padding after the last node may lie past the caller's arena
  CK(s.node_bytes.reserve(span + 16));
  CK(s.node_off.reserve(8 * n));
  CK(s.node_len.reserve(4 * n));
  CK(s.digests.reserve(32 * n));
  CK(s.order.reserve(4 * n));
  CK(s.bins.reserve(kBinScratchWords * sizeof(uint32_t)));
  if (avail < span + 16) CK(cudaMemsetAsync(s.node_bytes.as<uint8_t>() + avail, 0, span + 16 - avail, st));
  CK(cudaMemcpyAsync(s.node_bytes.p, node_bytes + lo, avail, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.node_off.p, node_off + i0, 8 * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.node_len.p, node_len + i0, 4 * n, cudaMemcpyHostToDevice, st));
  const uint32_t* ord = nullptr;
  if (ctx->binning) {
    CK(launch_bin_nodes(s.node_len.as<uint32_t>(), nullptr, n, s.bins.as<uint32_t>(), s.order.as<uint32_t>(), st, nullptr,
                        nullptr, ctx->long_leaf_bin));
    ord = s.order.as<uint32_t>();
  }
  // node_off values stay the caller's: byte_base = lo translates them
  CK(launch_keccak256_nodes(s.node_bytes.as<uint8_t>(), lo, s.node_off.as<uint64_t>(), s.node_len.as<uint32_t>(), ord, n,
                            s.digests.as<uint8_t>(), nullptr, s.bins.as<uint32_t>() + 2 * kNumBins, d.sm_count, st,
                            ord ? bin_split_word(s.bins.as<uint32_t>()) : nullptr, ctx->long_leaf_ctas));
  CK(cudaMemcpyAsync(digests32 + 32 * i0, s.digests.p, 32 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MPTV_OK;
}

// digest_keccak over a whole arena: the nodes are cut into contiguous index ranges with equal shares of the bytes,
// one per device of the context (host thread + stream each, no inter-device traffic)
static int keccak256_batch_run(mptv_ctx* ctx, const uint8_t* node_bytes, uint64_t node_bytes_len, const uint64_t* node_off,
                               const uint32_t* node_len, uint64_t n_nodes, uint8_t* digests32) {
  if (!ctx || (n_nodes && (!node_bytes || !node_off || !node_len || !digests32))) return MPTV_ERR_ARG;
  if (n_nodes == 0) return MPTV_OK;
  if (n_nodes > 0xfffffff0ull) return MPTV_ERR_ARG;
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_nodes; i++) {
    if (node_off[i] & 15) return MPTV_ERR_ALIGN;
    if (node_len[i] > kMaxNodeLen) return MPTV_ERR_ARG;
    if (node_off[i] > node_bytes_len || node_len[i] > node_bytes_len - node_off[i]) return MPTV_ERR_ARG;  // no wrap, exact end
    total += node_len[i];
  }
  const int nd = (n_nodes < 4096) ? 1 : (int)ctx->dev.size();  // a small call is not worth the threads
  std::vector<uint64_t> cut(nd + 1, 0);
  cut[nd] = n_nodes;
  if (nd > 1) {
    uint64_t acc = 0;
    int k = 1;
    for (uint64_t i = 0; i < n_nodes && k < nd; i++) {
      acc += node_len[i];
      while (k < nd && acc >= total / nd * k) cut[k++] = i + 1;
    }
    for (; k < nd; k++) cut[k] = n_nodes;
  }
  std::vector<int> rcs(nd, MPTV_OK);
  if (nd == 1) rcs[0] = keccak256_slice(ctx, ctx->dev[0], node_bytes, node_bytes_len, node_off, node_len, 0, n_nodes, digests32);
  else {
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] {
        rcs[k] = keccak256_slice(ctx, ctx->dev[k], node_bytes, node_bytes_len, node_off, node_len, cut[k], cut[k + 1], digests32);
      });
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nd; k++) if (rcs[k] != MPTV_OK) return rcs[k];
  return MPTV_OK;
}

int mptv_keccak256_batch(mptv_ctx* ctx, const uint8_t* node_bytes, uint64_t node_bytes_len, const uint64_t* node_off,
                         const uint32_t* node_len, uint64_t n_nodes, uint8_t* digests32) {
  int rc;
  try {
    rc = keccak256_batch_run(ctx, node_bytes, node_bytes_len, node_off, node_len, n_nodes, digests32);
  } catch (...) {
    rc = MPTV_ERR_NOMEM;
  }
  if (rc != MPTV_OK && ctx) for (Device& d : ctx->dev) quiesce(d);
  return rc;
}

}  // extern "C"
