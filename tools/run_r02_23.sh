#!/bin/bash
cd /root/repo
timeout 400 python -m pytest tests/test_gpu_borsh.py tests/test_gpu_storage_borsh.py tests/test_gpu_single.py -x -q -m gpu > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02p_tests.log
python bench.py --no-configs --no-cpu-baseline --steps 3 > gpurun_out/r02p_bench_quick.json 2> gpurun_out/r02p_bench_quick.err; echo "quick bench rc=$?"
python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02p_bench_quick.json') if l.startswith('{')][0]); e=j['e2e']; print(j['value'], e['value'], e['host_ms'])"
