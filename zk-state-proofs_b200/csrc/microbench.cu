// microbench.cu -- integer issue-rate probe: the roofline denominator of the Keccak kernel.
//
// P_int = alu-pipe lane-ops / s for LOP3 / SHF, "must be re-measured on B200
// with a LOP3/SHF microbenchmark".  Three instruction mixes: pure LOP3, pure SHF, and the Keccak
// round mix (122 LOP3 : 58 SHF).  Chains are independent enough (16 live values per thread, 8
// resident warps per scheduler) that the pipe, not latency, is the limit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mptv {

// MODE 0: LOP3   1: SHF   2: Keccak mix (11 LOP3 : 5 SHF)
//      3: IMAD (low 32)   4: IMAD.HI   5: IMAD.WIDE   6: LOP3 + IMAD 1:1   7: LOP3 + IMAD.HI 1:1
//      8: LOP3 + IMAD.WIDE 2:1
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* out, int iters) {
  uint32_t a[16];
#pragma unroll
  for (int j = 0; j < 16; j++) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t x = a[j], y = a[(j + 1) & 15], z = a[(j + 5) & 15];
      uint32_t r;
      bool fma_op = false;
      if (MODE == 6 || MODE == 7) fma_op = (j & 1);
      if (MODE == 8) fma_op = (j % 3) == 2;
      if (MODE == 1 || (MODE == 2 && (j % 3) == 2)) {
        asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(r) : "r"(y), "r"(z));
      } else if (MODE == 3 || (MODE == 6 && fma_op)) {
        asm volatile("mad.lo.u32 %0, %1, 128, %2;" : "=r"(r) : "r"(y), "r"(z));
      } else if (MODE == 4 || (MODE == 7 && fma_op)) {
        asm volatile("mad.hi.u32 %0, %1, 128, %2;" : "=r"(r) : "r"(y), "r"(z));
      } else if (MODE == 5 || (MODE == 8 && fma_op)) {
        uint64_t w;
        asm volatile("mad.wide.u32 %0, %1, 128, %2;" : "=l"(w) : "r"(y), "l"(((uint64_t)z << 32) | x));
        r = (uint32_t)w ^ (uint32_t)(w >> 32);
      } else {
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(y), "r"(z));
      }
      a[j] = r;
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) x ^= a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// returns lane-operations per second (warp instruction = 32 lane-ops) in *ops_per_s
cudaError_t run_int_peak(int mode, int sm_count, uint32_t* scratch /* >= sm_count*8*256 u32 */, cudaStream_t st,
                         double* ops_per_s) {
  const int blocks = sm_count * 8, iters = 8192;
  cudaEvent_t e0, e1;
  cudaError_t e = cudaEventCreate(&e0);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&e1);
  if (e != cudaSuccess) return e;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, st);
    switch (mode) {
      case 0: k_int_peak<0><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 1: k_int_peak<1><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 2: k_int_peak<2><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 3: k_int_peak<3><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 4: k_int_peak<4><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 5: k_int_peak<5><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 6: k_int_peak<6><<<blocks, 256, 0, st>>>(scratch, iters); break;
      case 7: k_int_peak<7><<<blocks, 256, 0, st>>>(scratch, iters); break;
      default: k_int_peak<8><<<blocks, 256, 0, st>>>(scratch, iters); break;
    }
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return e;
  *ops_per_s = (double)blocks * 256.0 * (double)iters * 16.0 / (best * 1e-3);
  return cudaGetLastError();
}

}  // namespace mptv
