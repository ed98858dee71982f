#!/bin/bash
# streamed borsh entry: parity tests, then the chunk-size x threads sweep on the config-2 batch
cd /root/repo
python -m pytest tests/test_gpu_borsh.py tests/test_gpu_errors.py -x -q -m gpu > gpurun_out/borsh_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/borsh_tests.log
python tools/borsh_stream_bench.py > gpurun_out/borsh_sweep.txt 2> gpurun_out/borsh_sweep.err; echo "sweep rc=$?"; cat gpurun_out/borsh_sweep.txt; tail -3 gpurun_out/borsh_sweep.err
