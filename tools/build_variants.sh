#!/bin/bash
# build libmptv variants with different Keccak round-loop unroll factors into build/variants/
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for u in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-pthread -shared -cudart static \
    -DMPTV_KECCAK_UNROLL=$u -o build/variants/libmptv_u$u.so \
    zk-state-proofs_b200/csrc/keccak_kernels.cu zk-state-proofs_b200/csrc/verify_kernels.cu \
    zk-state-proofs_b200/csrc/microbench.cu zk-state-proofs_b200/csrc/mptv_api.cu &
done
wait
ls -la build/variants
