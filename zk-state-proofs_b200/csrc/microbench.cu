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

template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* out, int iters) {
  uint32_t a[16];
#pragma unroll
  for (int j = 0; j < 16; j++) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const bool shf = MODE == 1 || (MODE == 2 && (j % 3) == 2);  // mode 2: 5 SHF : 11 LOP3 ~ 58:122
      if (shf) {
        asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(a[j]) : "r"(a[(j + 1) & 15]), "r"(a[(j + 5) & 15]));
      } else {
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[j]) : "r"(a[j]), "r"(a[(j + 1) & 15]), "r"(a[(j + 5) & 15]));
      }
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
    if (mode == 0) k_int_peak<0><<<blocks, 256, 0, st>>>(scratch, iters);
    else if (mode == 1) k_int_peak<1><<<blocks, 256, 0, st>>>(scratch, iters);
    else k_int_peak<2><<<blocks, 256, 0, st>>>(scratch, iters);
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
