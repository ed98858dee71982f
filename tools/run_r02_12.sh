#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_borsh.py -x -q -m gpu -k device_flatten > gpurun_out/r02_12_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_12_tests.log
python tools/borsh_stream_bench.py 1000000 quick 2>gpurun_out/r02_12.err | grep -E "host_dedup 1|device flatten|hybrid"; tail -3 gpurun_out/r02_12.err
