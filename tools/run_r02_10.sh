#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02_10_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_10_tests.log
for i in 1 2 3; do python tools/borsh_stream_bench.py 1000000 quick 2>/dev/null | grep "host_dedup 1"; done
python tools/flatten_micro_dump.py 1000000 10000000 > /dev/null && python tools/flatten_probe_sweep.py 2>/dev/null | grep "libmptv.so"
