#!/bin/bash
# round 2, session 9: compute-sanitizer on the new kernels (small cases), then a differential fuzz campaign
cd /root/repo
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_single.py -x -q -m gpu -k "limits or storage_guest" > gpurun_out/r02_sanitizer_single.log 2>&1; echo "memcheck single rc=$?"; tail -4 gpurun_out/r02_sanitizer_single.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_storage.py tests/test_gpu_borsh.py -x -q -m gpu -k "hashed_keys_on_the_device or bad_root" > gpurun_out/r02_sanitizer_keys_borsh.log 2>&1; echo "memcheck keys/borsh rc=$?"; tail -4 gpurun_out/r02_sanitizer_keys_borsh.log
timeout 1500 python tools/fuzz_gpu_vs_oracle.py 6 9000 > gpurun_out/r02_fuzz_gpu_vs_oracle.txt 2>&1; echo "fuzz rc=$?"; tail -4 gpurun_out/r02_fuzz_gpu_vs_oracle.txt
