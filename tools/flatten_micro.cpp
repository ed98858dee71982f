// flatten_micro.cpp -- where the host stage of mptv_verify_borsh spends its time: the node walk of
// host_flatten.h with parts switched off, on a blob file dumped by tools/flatten_micro_dump.py.
//   g++ -O2 -mavx2 -std=c++17 -pthread -Izk-state-proofs_b200/csrc -I/usr/local/cuda/include tools/flatten_micro.cpp \
//       zk-state-proofs_b200/csrc/host_flatten.cpp -o build/flatten_micro
//   build/flatten_micro /tmp/blobs.bin /tmp/boff.bin
// mode bits: 1 fingerprint  2 table lookup (with memcmp)  4 copy unique nodes (NT stores)  8 lookup WITHOUT memcmp
//            16 stream prefetch (T0)  32 stream prefetch (NTA)  64 memcmp only against the previous copy (no table)
//            128 copy with ordinary stores  256 copy with NT stores, nodes on 64-byte boundaries, whole lines
#include <immintrin.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "host_flatten.h"
using namespace mptv;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t nb = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> blobs(nb + 8192); if (fread(blobs.data(), 1, nb, f)) {} fclose(f);
  f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); size_t no = ftell(f) / 8; fseek(f, 0, SEEK_SET);
  std::vector<uint64_t> boff(no); if (fread(boff.data(), 8, no, f)) {} fclose(f);
  const uint64_t n = no - 1;
  const size_t chunk = 32u << 20;
  const unsigned hw = std::thread::hardware_concurrency();
  // chunk boundaries
  std::vector<uint64_t> cb = {0};
  for (uint64_t i = 1; i <= n; i++) if (boff[i] - boff[cb.back()] > chunk || i == n) cb.push_back(i);
  printf("%llu blobs, %.2f GB, %zu chunks of 32 MB, %u cores\n", (unsigned long long)n, nb / 1e9, cb.size() - 1, hw);
  DedupTable tab; tab.reserve(256 << 10);
  const int modes[] = {16, 1 | 2 | 16, 4 | 16, 128 | 16, 256 | 16, 1 | 2 | 4 | 16, 1 | 2 | 128 | 16, 1 | 2 | 256 | 16};
  for (unsigned T : {1u, 4u, hw}) {
    std::vector<std::vector<uint8_t>> outv(T);
    std::vector<uint8_t*> out(T);
    for (unsigned t = 0; t < T; t++) { outv[t].resize((chunk / T + (1 << 20)) * 3 + 64); out[t] = (uint8_t*)(((uintptr_t)outv[t].data() + 63) & ~(uintptr_t)63); }
    for (int mode : modes) {
      double best = 1e9; uint64_t dups = 0, nodes = 0;
      for (int rep = 0; rep < 3; rep++) {
        std::atomic<uint64_t> a_dups(0), a_nodes(0);
        double t0 = now();
        for (size_t c = 0; c + 1 < cb.size(); c++) {
          tab.new_epoch();
          const uint64_t cs = cb[c], ce = cb[c + 1], per = (ce - cs + T - 1) / T;
          std::vector<std::thread> th;
          auto work = [&](unsigned t) {
            const uint64_t lo = std::min(ce, cs + per * t), hi = std::min(ce, lo + per);
            size_t at = 0; uint64_t d = 0, nn = 0; uint64_t acc = 0;
            const uint8_t* prev[16] = {nullptr}; uint32_t prevlen[16] = {0};
            for (uint64_t i = lo; i < hi; i++) {
              const uint8_t* p = blobs.data() + boff[i];
              uint32_t cnt = rd_u32(p); p += 4;
              uint64_t fp = 0;
              if (mode & 1) { uint32_t l0 = rd_u32(p); if (l0 >= 128) { fp = node_fingerprint(p + 4, l0); if (mode & 10) tab.prefetch(fp); } }
              for (uint32_t j = 0; j < cnt; j++) {
                uint32_t len = rd_u32(p);
                if (mode & 16) for (uint32_t pf = 0; pf < len + 4; pf += 64) __builtin_prefetch(p + 2048 + pf);
                if (mode & 32) for (uint32_t pf = 0; pf < len + 4; pf += 64) __builtin_prefetch(p + 2048 + pf, 0, 0);
                uint64_t fpn = 0;
                if ((mode & 1) && j + 1 < cnt) { const uint8_t* q = p + 4 + len; uint32_t l1 = rd_u32(q); if (l1 >= 128) { fpn = node_fingerprint(q + 4, l1); if (mode & 10) tab.prefetch(fpn); } }
                bool dup = false;
                if ((mode & 2) && len >= 128) { uint32_t o16; dup = tab.find_or_insert(p + 4, len, fp, (uint32_t)(at >> 4), &o16); }
                if ((mode & 8) && len >= 128) { uint32_t o16; dup = tab.find_or_insert(p + 4, 0, fp, (uint32_t)(at >> 4), &o16); }
                if ((mode & 64) && j < 16) { dup = prevlen[j] == len && memcmp(prev[j], p + 4, len) == 0; prev[j] = p + 4; prevlen[j] = len; }
                if (dup) d++;
                else if (mode & 256) {
                  const uint32_t full = (len + 63) & ~63u;
                  for (uint32_t q = 0; q < full; q += 32)
                    _mm256_stream_si256((__m256i*)(out[t] + at + q), _mm256_loadu_si256((const __m256i*)(p + 4 + q)));
                  at += full;
                } else if (mode & 128) { memcpy(out[t] + at, p + 4, len); at += up16z(len); }
                else { if (mode & 4) copy_node_stream(out[t] + at, p + 4, len); at += up16z(len); }
                acc += fp; fp = fpn;
                p += 4 + len; nn++;
              }
            }
            a_dups += d; a_nodes += nn + (acc == 12345);
          };
          for (unsigned t = 1; t < T; t++) th.emplace_back(work, t);
          work(0);
          for (auto& x : th) x.join();
        }
        best = std::min(best, now() - t0); dups = a_dups; nodes = a_nodes;
      }
      printf("T=%2u mode %3d [%s%s%s%s%s%s%s]: %7.1f ms  %6.1f ns/node/thread  %5.1f GB/s read  dups %llu/%llu\n", T, mode,
             mode & 1 ? "fp " : "", mode & 2 ? "lookup+memcmp " : "", mode & 8 ? "lookup-nocmp " : "", mode & 64 ? "cmp-prev " : "",
             mode & 4 ? "copy " : (mode & 128 ? "copy-plain " : (mode & 256 ? "copy-nt64 " : "")), mode & 16 ? "pf " : "", mode & 32 ? "pfnta " : "", best * 1e3, best * 1e9 * T / nodes, nb / best / 1e9,
             (unsigned long long)dups, (unsigned long long)nodes);
      fflush(stdout);
    }
  }
}
