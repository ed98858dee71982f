// keccak_f1600.cuh -- Keccak-f[1600] with the state held in registers as 25 x (lo,hi) u32 pairs.
//
// Replaces tiny_keccak::keccakf (tiny-keccak 2.0.2, Cargo.lock:7879-7894; reached from
// crypto_ops::keccak::digest_keccak, /root/reference/crypto-ops/src/keccak.rs:6-12).
//
// Instruction budget per round: theta 20 LOP3 (column parities, 3-input
// XOR) + 10 SHF (rot-1 of the parities) + 50 LOP3 (A ^= C[x-1] ^ rot(C[x+1]) as ONE 3-input
// LOP3 per half -- written as inline lop3 so the compiler cannot re-associate it into a
// separate D[x] plus a 2-input XOR); rho 48 SHF (funnel shifts, pi is register renaming);
// chi 50 LOP3 (a ^ (~b & c), LUT 0xD2); iota 2 LOP3.  180 alu-pipe instructions / round,
// 4320 / permutation.
#pragma once
#include <stdint.h>

namespace mptv {

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a ^ (~b & c)
__device__ __forceinline__ uint32_t lop3_chi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// 64-bit rotate-left by a compile-time amount on a (lo,hi) pair: two SHF.L.W funnel shifts.
template <int N>
__device__ __forceinline__ void rotl64(uint32_t lo, uint32_t hi, uint32_t& olo, uint32_t& ohi) {
  if (N == 0) { olo = lo; ohi = hi; }
  else if (N == 32) { olo = hi; ohi = lo; }
  else if (N < 32) {
    ohi = __funnelshift_l(lo, hi, (uint32_t)N);
    olo = __funnelshift_l(hi, lo, (uint32_t)N);
  } else {
    ohi = __funnelshift_l(hi, lo, (uint32_t)(N - 32));
    olo = __funnelshift_l(lo, hi, (uint32_t)(N - 32));
  }
}

__device__ __constant__ uint32_t kRC[48] = {  // (lo,hi) of the 24 iota constants
  0x00000001u, 0x00000000u, 0x00008082u, 0x00000000u, 0x0000808au, 0x80000000u, 0x80008000u, 0x80000000u,
  0x0000808bu, 0x00000000u, 0x80000001u, 0x00000000u, 0x80008081u, 0x80000000u, 0x00008009u, 0x80000000u,
  0x0000008au, 0x00000000u, 0x00000088u, 0x00000000u, 0x80008009u, 0x00000000u, 0x8000000au, 0x00000000u,
  0x8000808bu, 0x00000000u, 0x0000008bu, 0x80000000u, 0x00008089u, 0x80000000u, 0x00008003u, 0x80000000u,
  0x00008002u, 0x80000000u, 0x00000080u, 0x80000000u, 0x0000800au, 0x00000000u, 0x8000000au, 0x80000000u,
  0x80008081u, 0x80000000u, 0x00008080u, 0x80000000u, 0x80000001u, 0x00000000u, 0x80008008u, 0x80000000u};

// One round: theta, rho+pi (into B), chi (back into A), iota.  Lane index = x + 5*y.
// rho offsets r[x+5y]; pi: B[y + 5*((2x+3y)%5)] = rot(A[x+5y], r[x+5y]).
#define MPTV_RHOPI(SRC, DST, R)                                            \
  rotl64<R>(lop3_xor3(lo[SRC], Cm_lo[(SRC) % 5], Cr_lo[(SRC) % 5]),        \
            lop3_xor3(hi[SRC], Cm_hi[(SRC) % 5], Cr_hi[(SRC) % 5]), blo[DST], bhi[DST]);

__device__ __forceinline__ void keccak_round(uint32_t (&lo)[25], uint32_t (&hi)[25], uint32_t rc_lo,
                                             uint32_t rc_hi) {
  uint32_t Cl[5], Ch[5];
#pragma unroll
  for (int x = 0; x < 5; x++) {
    Cl[x] = lop3_xor3(lop3_xor3(lo[x], lo[x + 5], lo[x + 10]), lo[x + 15], lo[x + 20]);
    Ch[x] = lop3_xor3(lop3_xor3(hi[x], hi[x + 5], hi[x + 10]), hi[x + 15], hi[x + 20]);
  }
  // Cm[x] = C[x-1], Cr[x] = rot(C[x+1], 1)
  uint32_t Cm_lo[5], Cm_hi[5], Cr_lo[5], Cr_hi[5];
#pragma unroll
  for (int x = 0; x < 5; x++) {
    Cm_lo[x] = Cl[(x + 4) % 5];
    Cm_hi[x] = Ch[(x + 4) % 5];
    rotl64<1>(Cl[(x + 1) % 5], Ch[(x + 1) % 5], Cr_lo[x], Cr_hi[x]);
  }
  uint32_t blo[25], bhi[25];
  MPTV_RHOPI(0, 0, 0)    MPTV_RHOPI(1, 10, 1)   MPTV_RHOPI(2, 20, 62)  MPTV_RHOPI(3, 5, 28)   MPTV_RHOPI(4, 15, 27)
  MPTV_RHOPI(5, 16, 36)  MPTV_RHOPI(6, 1, 44)   MPTV_RHOPI(7, 11, 6)   MPTV_RHOPI(8, 21, 55)  MPTV_RHOPI(9, 6, 20)
  MPTV_RHOPI(10, 7, 3)   MPTV_RHOPI(11, 17, 10) MPTV_RHOPI(12, 2, 43)  MPTV_RHOPI(13, 12, 25) MPTV_RHOPI(14, 22, 39)
  MPTV_RHOPI(15, 23, 41) MPTV_RHOPI(16, 8, 45)  MPTV_RHOPI(17, 18, 15) MPTV_RHOPI(18, 3, 21)  MPTV_RHOPI(19, 13, 8)
  MPTV_RHOPI(20, 14, 18) MPTV_RHOPI(21, 24, 2)  MPTV_RHOPI(22, 9, 61)  MPTV_RHOPI(23, 19, 56) MPTV_RHOPI(24, 4, 14)
#pragma unroll
  for (int y = 0; y < 25; y += 5) {
#pragma unroll
    for (int x = 0; x < 5; x++) {
      lo[y + x] = lop3_chi(blo[y + x], blo[y + (x + 1) % 5], blo[y + (x + 2) % 5]);
      hi[y + x] = lop3_chi(bhi[y + x], bhi[y + (x + 1) % 5], bhi[y + (x + 2) % 5]);
    }
  }
  lo[0] ^= rc_lo;
  hi[0] ^= rc_hi;
}
#undef MPTV_RHOPI

#ifndef MPTV_KECCAK_UNROLL
#define MPTV_KECCAK_UNROLL 6
#endif

constexpr int kKeccakUnroll = MPTV_KECCAK_UNROLL;

__device__ __forceinline__ void keccak_f1600(uint32_t (&lo)[25], uint32_t (&hi)[25]) {
#pragma unroll kKeccakUnroll
  for (int r = 0; r < 24; r++) keccak_round(lo, hi, kRC[2 * r], kRC[2 * r + 1]);
}

}  // namespace mptv
