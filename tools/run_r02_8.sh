#!/bin/bash
# round 2, session 8: pull mode of the borsh stream -- parity, then the sweep with page-locked blobs
cd /root/repo
python -m pytest tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02_8_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_8_tests.log
python tools/borsh_stream_bench.py 1000000 quick pinned 2>gpurun_out/r02_8.err | tee gpurun_out/r02_borsh_pull_quick.txt; tail -3 gpurun_out/r02_8.err
