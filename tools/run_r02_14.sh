#!/bin/bash
# final default bench line on the tree with the no-idle-gap warm-up, then a fuzz campaign that includes the device-flatten modes
cd /root/repo
python bench.py > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02d_bench_default.err
timeout 420 python tools/fuzz_gpu_vs_oracle.py 12 9000 > gpurun_out/r02d_fuzz.txt 2>&1; echo "fuzz rc=$?"; tail -4 gpurun_out/r02d_fuzz.txt
