#!/bin/bash
# 2 GPUs: the multi-GPU tests on the final tree (incl. the storage stream over two devices)
cd /root/repo
python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02i_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -15 gpurun_out/r02i_multi_tests.log
