#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_single.py tests/test_gpu_storage.py -x -q -m gpu > gpurun_out/r02_11_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_11_tests.log
MPTV_LIB=build/variants/libmptv_smalltiming.so python tools/latency_probe.py | tee gpurun_out/r02_latency_probe.txt
python tools/fuzz_gpu_vs_oracle.py 40 12000 > gpurun_out/r02_fuzz_gpu_vs_oracle.txt 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r02_fuzz_gpu_vs_oracle.txt
