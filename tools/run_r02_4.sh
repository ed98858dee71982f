#!/bin/bash
# round 2, session 4: latency path -- parity tests, config 1 latency
cd /root/repo
python -m pytest tests/test_gpu_single.py tests/test_gpu_parity.py tests/test_gpu_storage.py tests/test_gpu_errors.py -x -q -m gpu > gpurun_out/r02_4_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_4_tests.log
python bench.py --workload config1 --steps 5 --warmup 3 > gpurun_out/r02_bench_config1.json 2> gpurun_out/r02_4_c1.err; echo "config1 rc=$?"; tail -2 gpurun_out/r02_4_c1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_config1.json')); print('latency_us', d['latency_us'], 'launches/call', d['gpu_launches']/(5*200), 'cpu', d['cpu_baseline']['latency_us'])"
