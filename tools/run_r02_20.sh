#!/bin/bash
cd /root/repo
python -m pytest tests/test_cpp_host.py tests/test_gpu_storage.py tests/test_cabi.py -q -m gpu > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02l_tests.log
