#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_storage.py tests/test_gpu_single.py tests/test_gpu_rebuild.py tests/test_gpu_errors.py tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02_7_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_7_tests.log
