#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_borsh.py -x -q -m gpu -k write_combining > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02k_tests.log
python tools/wc_probe.py > gpurun_out/r02k_wc_probe.txt 2> gpurun_out/r02k_wc_probe.err; echo "probe rc=$?"; cat gpurun_out/r02k_wc_probe.txt; tail -3 gpurun_out/r02k_wc_probe.err
