#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dedup" > gpurun_out/dedup2_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/dedup2_tests.log
python bench.py --dedup --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/dedup2_bench.json 2> gpurun_out/dedup2_bench.err; echo "bench rc=$?"
