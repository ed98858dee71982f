// host_codec.h -- internal: the borsh(MerkleProofInput) walkers of host_codec.cpp, shared with the streamed
// host entry in mptv_api.cu (wire format: /root/reference/crypto-ops/src/types.rs:4-9, u32-LE length prefixes).
#pragma once
#include <emmintrin.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mptv {

struct BlobShape { uint32_t n_nodes; uint32_t key_len; uint64_t padded_bytes; uint8_t ok; uint8_t bad_root; };

inline bool borsh_u32(const uint8_t* p, const uint8_t* end, uint32_t& v) {
  if (end - p < 4) return false;
  v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  return true;
}

// pass 1: walk one blob, no copies; ok = well-formed and nothing left over (borsh::from_slice rejects trailing bytes)
inline BlobShape borsh_shape(const uint8_t* p, const uint8_t* end) {
  BlobShape s = {0, 0, 0, 0, 0};
  uint32_t n, len;
  if (!borsh_u32(p, end, n)) return s;
  p += 4;
  for (uint32_t i = 0; i < n; i++) {
    if (!borsh_u32(p, end, len) || (uint64_t)(end - p - 4) < len) return s;
    p += 4 + len;
    s.padded_bytes += ((uint64_t)len + 15) & ~15ull;
  }
  s.n_nodes = n;
  if (!borsh_u32(p, end, len) || (uint64_t)(end - p - 4) < len) return s;
  s.bad_root = len != 32;
  p += 4 + len;
  if (!borsh_u32(p, end, len) || (uint64_t)(end - p - 4) < len) return s;
  s.key_len = len;
  p += 4 + len;
  s.ok = p == end;
  return s;
}

// One node into the 16-byte aligned arena with non-temporal stores: the arena is written once and next read by
// the DMA engine, so going around the cache saves the read-for-ownership of every destination line (a third of
// the flattener's memory traffic).  Writes whole 16-byte chunks, i.e. the zero padding too.
inline void copy_node_stream(uint8_t* dst, const uint8_t* src, uint32_t len) {
  const uint32_t full = len & ~15u;
  for (uint32_t i = 0; i < full; i += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
  if (len & 15u) {
    alignas(16) uint8_t tail[16] = {0};
    memcpy(tail, src + full, len & 15u);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + full), _mm_load_si128(reinterpret_cast<const __m128i*>(tail)));
  }
}

// pass 2: copy one well-formed blob into the CSR arrays.  Nodes go to node_bytes[o ...] (16-byte aligned, padding
// zeroed) as nodes k, k+1, ...; node_src (optional) receives each node's byte position relative to `src_base`.
inline void borsh_copy(const uint8_t* blob, const BlobShape& sh, uint8_t* node_bytes, uint64_t o, uint64_t* node_off,
                       uint32_t* node_len, uint64_t k, uint8_t* root32, uint8_t* key_dst, uint64_t* node_src,
                       const uint8_t* src_base) {
  const uint8_t* p = blob + 4;
  for (uint32_t j = 0; j < sh.n_nodes; j++) {
    const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    copy_node_stream(node_bytes + o, p + 4, len);  // o is a multiple of 16; callers fence before publishing
    const uint64_t pad = (((uint64_t)len + 15) & ~15ull) - len;
    node_off[k] = o; node_len[k] = len;
    if (node_src) node_src[k] = (uint64_t)(p + 4 - src_base);
    k++; o += len + pad; p += 4 + len;
  }
  const uint32_t rl = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  if (rl == 32) memcpy(root32, p + 4, 32); else memset(root32, 0, 32);
  p += 4 + rl;
  if (sh.key_len) memcpy(key_dst, p + 4, sh.key_len);
}

template <class F>
void parallel_for(uint64_t n, int n_threads, F f) {
  if (n_threads <= 1 || n < 1024) { f(0, n); return; }
  std::vector<std::thread> th;
  const uint64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const uint64_t lo = std::min(n, per * t), hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& x : th) x.join();
}

// A few worker threads that live for one call: run(f) executes f(tid) on all of them (tid 0 = the caller) and
// returns when every one is done; barrier() may be used inside f (all threads must reach it).
// The streamed entries call run() a hundred times per call, a chunk every few hundred microseconds, so a handoff
// through a condition variable (tens of microseconds of wake-up latency, twice per chunk) would cost a sixth of the
// call.  Workers therefore SPIN for the next job for a short while and only then go to sleep; the caller does the same
// for completion.  Every wait is bounded spinning followed by yield / sleep, so the pool also behaves when there are
// more threads than cores (a preempted worker is not waited for by fifteen spinning ones).
class WorkerPool {
 public:
  explicit WorkerPool(int n) : n_(n < 1 ? 1 : n) {
    for (int t = 1; t < n_; t++) th_.emplace_back([this, t] { loop(t); });
  }
  ~WorkerPool() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; gen_.fetch_add(1, std::memory_order_release); }
    cv_.notify_all();
    for (auto& x : th_) x.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& f) {
    job_ = &f;
    pending_.store(n_ - 1, std::memory_order_relaxed);
    {
      // the generation is bumped under the mutex so that a worker that decided to sleep cannot miss it
      std::lock_guard<std::mutex> g(mu_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    f(0);
    for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; spins++) {
      if (spins < kSpin) _mm_pause(); else std::this_thread::yield();
    }
  }
  void barrier() {
    const int g = bar_gen_.load(std::memory_order_acquire);
    if (bar_count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
      bar_count_.store(0, std::memory_order_relaxed);
      bar_gen_.fetch_add(1, std::memory_order_release);
    } else {
      for (int spins = 0; bar_gen_.load(std::memory_order_acquire) == g; spins++) {
        if (spins < kSpin) _mm_pause(); else std::this_thread::yield();
      }
    }
  }

 private:
  static constexpr int kSpin = 20000;  // ~ 50-100 us of pause instructions before giving the core away
  void loop(int t) {
    uint64_t last = 0;
    for (;;) {
      // wait for the next generation: spin first, then sleep
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == last) {
        if (++spins < kSpin) { _mm_pause(); continue; }
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_.fetch_add(1, std::memory_order_acq_rel);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != last; });
        sleepers_.fetch_sub(1, std::memory_order_acq_rel);
      }
      last = gen_.load(std::memory_order_acquire);
      if (stop_) return;
      (*job_)(t);
      pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  const std::function<void(int)>* job_ = nullptr;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
  std::atomic<int> bar_count_{0}, bar_gen_{0};
};

}  // namespace mptv
