// host_flatten.cpp -- the candidate table of the de-duplicating chunk builder (host_flatten.h).  Host code only; no
// hashing of nodes in the Keccak sense, no RLP decoding, no trie walking: layout work in front of the GPU.
#include "host_flatten.h"

#include <stdlib.h>
#include <sys/mman.h>

namespace mptv {

void DedupTable::release() {
  free(tab_);
  tab_ = nullptr;
  mask_ = 0;
}

bool DedupTable::reserve(size_t entries) {
  size_t cap = 1 << 12;
  while (cap < entries) cap <<= 1;
  if (tab_ && cap <= mask_ + 1) return true;
  release();
  void* p = nullptr;
  // probes land on random lines of the table: on 2 MB pages (transparent huge pages, where the host allows madvise)
  // they do not also miss the TLB
  const size_t bytes = cap * sizeof(DedupEntry);
  if (posix_memalign(&p, bytes >= (2u << 20) ? (2u << 20) : 64, bytes) != 0) return false;
#ifdef MADV_HUGEPAGE
  if (bytes >= (2u << 20)) madvise(p, bytes, MADV_HUGEPAGE);
#endif
  memset(p, 0, bytes);
  tab_ = static_cast<DedupEntry*>(p);
  mask_ = cap - 1;
  epoch_ = 0;
  return true;
}

void DedupTable::new_epoch() {
  if (!tab_) return;
  if (++epoch_ > 0xffffull) {  // the tag wrapped: really clear
    memset(static_cast<void*>(tab_), 0, (mask_ + 1) * sizeof(DedupEntry));
    epoch_ = 1;
  }
}

// exact compare of two nodes of len >= 32 bytes: no early exit (candidates with equal fingerprints are almost always
// equal), 32 bytes a step, the tail as an overlapping last step
static inline bool equal_bytes(const uint8_t* a, const uint8_t* b, uint32_t len) {
  __m256i acc = _mm256_setzero_si256();
  uint32_t i = 0;
  for (; i + 32 <= len; i += 32)
    acc = _mm256_or_si256(acc, _mm256_xor_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + i)),
                                                _mm256_loadu_si256(reinterpret_cast<const __m256i*>(b + i))));
  acc = _mm256_or_si256(acc, _mm256_xor_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + len - 32)),
                                              _mm256_loadu_si256(reinterpret_cast<const __m256i*>(b + len - 32))));
  return _mm256_testz_si256(acc, acc) != 0;
}

bool DedupTable::find_or_insert(const uint8_t* p, uint32_t len, uint64_t h, uint32_t my_off16, uint32_t* off16) {
  if (!tab_ || !epoch_) return false;
  const uint64_t tag = epoch_ << 48;
  const uint64_t want = tag | (h >> 16);
  size_t slot = (size_t)h & mask_;
  for (int probe = 0; probe < 8; probe++, slot = (slot + 1) & mask_) {
    DedupEntry& e = tab_[slot];
    uint64_t k = e.key.load(std::memory_order_acquire);
    if ((k >> 48) != epoch_) {
      // free (never used, or left over from an earlier chunk): claim it, then publish src / len / offset
      if (e.key.compare_exchange_strong(k, want, std::memory_order_acq_rel)) {
        e.src = p;
        e.len = len;
        e.ready.store(tag | my_off16, std::memory_order_release);
        return false;
      }
      // lost the race for this entry: k now holds the winner's key
    }
    if (k != want) continue;
    uint64_t r;
    while (((r = e.ready.load(std::memory_order_acquire)) >> 48) != epoch_) _mm_pause();  // the winner is between its two stores
    if (e.len == len && equal_bytes(e.src, p, len)) {
      *off16 = (uint32_t)r;
      return true;
    }
    // same fingerprint, different bytes: keep probing
  }
  return false;
}

}  // namespace mptv
