#!/bin/bash
# round 2, session 3: parity with unlimited inline nesting + borsh stream with the bounce-buffer writer
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_borsh.py tests/test_gpu_errors.py -x -q -m gpu > gpurun_out/r02_3_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_3_tests.log
python tools/borsh_stream_bench.py 1000000 > gpurun_out/r02_borsh_stream_quick.txt 2> gpurun_out/r02_3_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02_borsh_stream_quick.txt; tail -3 gpurun_out/r02_3_sweep.err
