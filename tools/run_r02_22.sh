#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_borsh.py -x -q -m gpu -k "parser_accepts" > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02o_tests.log
