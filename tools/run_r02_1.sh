#!/bin/bash
# round 2, session 1: borsh stream with host-side aliasing -- parity tests, sweep, bench
cd /root/repo
python -m pytest tests/test_gpu_borsh.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_1_tests.log
python tools/borsh_stream_bench.py > gpurun_out/r02_borsh_stream_sweep.txt 2> gpurun_out/r02_1_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02_borsh_stream_sweep.txt; tail -3 gpurun_out/r02_1_sweep.err
