#!/bin/bash
# round 2, session 2: borsh stream after the plain-store / look-ahead changes
cd /root/repo
python -m pytest tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02_2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_2_tests.log
python tools/borsh_stream_bench.py 1000000 quick > gpurun_out/r02_borsh_stream_quick.txt 2> gpurun_out/r02_2_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02_borsh_stream_quick.txt; tail -3 gpurun_out/r02_2_sweep.err
