#!/bin/bash
cd /root/repo
timeout 200 python -m pytest tests/test_cabi.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02q_tests.log
