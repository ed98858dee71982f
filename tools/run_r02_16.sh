#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_storage_borsh.py tests/test_gpu_storage.py tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r02g_tests.log
