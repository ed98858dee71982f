// CPU-only checks of the host helpers behind mptv_flatten_borsh / mptv_verify_borsh
// (zk-state-proofs_b200/csrc/host_codec.h): the worker pool + barrier, the non-temporal node copy and the
// borsh walkers.  Built by tests/test_host_codec.py, optionally under -fsanitize=thread.
#include <stdio.h>

#include <numeric>

#include "../../zk-state-proofs_b200/csrc/host_codec.h"

static int g_fail = 0;
#define CHECK(x) do { if (!(x)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); g_fail++; } } while (0)

static void test_pool_and_barrier() {
  for (int T : {1, 2, 5, 8}) {
    mptv::WorkerPool pool(T);
    std::vector<uint64_t> part(T), seen(T);
    uint64_t shared_total = 0;
    for (int round = 0; round < 300; round++) {
      pool.run([&](int t) {
        part[t] = (uint64_t)(round + 1) * (t + 1);
        pool.barrier();
        if (t == 0) shared_total = std::accumulate(part.begin(), part.end(), (uint64_t)0);  // written between the barriers
        pool.barrier();
        seen[t] = shared_total;
      });
      const uint64_t want = (uint64_t)(round + 1) * T * (T + 1) / 2;
      for (int t = 0; t < T; t++) CHECK(seen[t] == want);
    }
  }
}

static void test_copy_node_stream() {
  alignas(16) uint8_t dst[16 + 320 + 16];
  uint8_t src[400];
  for (int i = 0; i < 400; i++) src[i] = (uint8_t)(i * 7 + 1);
  for (uint32_t len = 0; len <= 300; len++)
    for (int mis = 0; mis < 5; mis++) {
      memset(dst, 0xee, sizeof dst);
      mptv::copy_node_stream(dst + 16, src + mis, len);
      _mm_sfence();
      const uint32_t padded = (len + 15) & ~15u;
      bool ok = memcmp(dst + 16, src + mis, len) == 0;
      for (uint32_t i = len; i < padded; i++) ok = ok && dst[16 + i] == 0;      // padding zeroed
      for (int i = 0; i < 16; i++) ok = ok && dst[i] == 0xee && dst[16 + padded + i] == 0xee;  // nothing outside
      CHECK(ok);
    }
}

static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }

static void test_borsh_walkers() {
  // proof [ 3 bytes, 0 bytes, 40 bytes ], root 32 bytes, key 2 bytes
  std::vector<uint8_t> b;
  put32(b, 3);
  put32(b, 3); b.insert(b.end(), {1, 2, 3});
  put32(b, 0);
  put32(b, 40); for (int i = 0; i < 40; i++) b.push_back((uint8_t)(100 + i));
  put32(b, 32); for (int i = 0; i < 32; i++) b.push_back((uint8_t)i);
  put32(b, 2); b.insert(b.end(), {0xab, 0xcd});
  mptv::BlobShape sh = mptv::borsh_shape(b.data(), b.data() + b.size());
  CHECK(sh.ok && !sh.bad_root && sh.n_nodes == 3 && sh.key_len == 2 && sh.padded_bytes == 16 + 0 + 48);
  for (size_t cut = 0; cut < b.size(); cut++) CHECK(!mptv::borsh_shape(b.data(), b.data() + cut).ok);  // truncated
  std::vector<uint8_t> more = b; more.push_back(0);
  CHECK(!mptv::borsh_shape(more.data(), more.data() + more.size()).ok);  // trailing byte
  alignas(16) uint8_t arena[16 + 64 + 16];
  memset(arena, 0xee, sizeof arena);
  uint64_t off[3], src[3];
  uint32_t len[3];
  uint8_t root[32], key[2];
  mptv::borsh_copy(b.data(), sh, arena, 16, off, len, 0, root, key, src, b.data());
  _mm_sfence();
  CHECK(off[0] == 16 && off[1] == 32 && off[2] == 32 && len[0] == 3 && len[1] == 0 && len[2] == 40);
  CHECK(arena[16] == 1 && arena[18] == 3 && arena[19] == 0 && arena[31] == 0 && arena[32] == 100 && arena[71] == 139);
  CHECK(arena[72] == 0 && arena[79] == 0 && arena[80] == 0xee && arena[15] == 0xee);
  CHECK(src[0] == 8 && src[1] == 15 && src[2] == 19 && root[31] == 31 && key[0] == 0xab && key[1] == 0xcd);
  std::vector<uint8_t> bad_root = b;
  bad_root[4 + 7 + 4 + 44] = 31;  // the root's length prefix
  CHECK(!mptv::borsh_shape(bad_root.data(), bad_root.data() + bad_root.size()).ok);  // lengths no longer add up
}

int main() {
  test_pool_and_barrier();
  test_copy_node_stream();
  test_borsh_walkers();
  printf("%d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
