#!/bin/bash
# overlapped pipeline: parity first, then the sweep (batch built once)
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors.py -x -q -m gpu > gpurun_out/ov_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ov_tests.log
python tools/sweep_overlap.py ${1:-config2} 5 > gpurun_out/ov_sweep_${1:-config2}.txt 2> gpurun_out/ov_sweep_${1:-config2}.err; echo "sweep rc=$?"
cat gpurun_out/ov_sweep_${1:-config2}.txt
