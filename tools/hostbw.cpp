// hostbw.cpp -- host memory bandwidth of the box as the flattener sees it: T threads each streaming its own
// slice (read-only sum, and read + non-temporal write), first-touch placed by the thread that uses it.
//   g++ -O2 -mavx2 -pthread tools/hostbw.cpp -o build/hostbw && build/hostbw [MiB per thread]
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const size_t mib = argc > 1 ? atoll(argv[1]) : 256;
  const size_t n = mib << 20;
  const unsigned hw = std::thread::hardware_concurrency();
  printf("hardware_concurrency %u, %zu MiB per thread\n", hw, mib);
  for (unsigned T : {1u, 2u, 4u, 8u, 12u, 16u, 24u, 32u, 48u, 64u}) {
    if (T > hw) break;
    std::vector<uint8_t*> src(T), dst(T);
    std::vector<std::thread> th;
    std::atomic<int> ready(0), go(0), phase2(0), done1(0);
    std::vector<uint64_t> sums(T);
    double t_read = 0, t_copy = 0;
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (unsigned t = 0; t < T; t++)
      th.emplace_back([&, t] {
        src[t] = (uint8_t*)aligned_alloc(64, n);
        dst[t] = (uint8_t*)aligned_alloc(64, n);
        memset(src[t], (int)t + 1, n);
        memset(dst[t], 0, n);
        ready++;
        while (!go.load()) _mm_pause();
        __m256i acc = _mm256_setzero_si256();
        for (int rep = 0; rep < 2; rep++)
          for (size_t i = 0; i < n; i += 32) acc = _mm256_add_epi64(acc, _mm256_load_si256((const __m256i*)(src[t] + i)));
        uint64_t o[4];
        _mm256_storeu_si256((__m256i*)o, acc);
        sums[t] = o[0] + o[1] + o[2] + o[3];
        done1++;
        while (!phase2.load()) _mm_pause();
        for (int rep = 0; rep < 2; rep++)
          for (size_t i = 0; i < n; i += 32)
            _mm256_stream_si256((__m256i*)(dst[t] + i), _mm256_load_si256((const __m256i*)(src[t] + i)));
        _mm_sfence();
      });
    while (ready.load() < (int)T) _mm_pause();
    t0 = now();
    go = 1;
    while (done1.load() < (int)T) _mm_pause();
    t1 = now();
    t2 = now();
    phase2 = 1;
    for (auto& x : th) x.join();
    t3 = now();
    t_read = t1 - t0; t_copy = t3 - t2;
    printf("T=%2u  read %7.1f GB/s   read+NT-write copy %7.1f GB/s (payload; x2 traffic)  [%llu]\n", T,
           2.0 * n * T / t_read / 1e9, 2.0 * n * T / t_copy / 1e9, (unsigned long long)sums[0]);
    for (unsigned t = 0; t < T; t++) { free(src[t]); free(dst[t]); }
  }
  return 0;
}
