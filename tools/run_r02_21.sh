#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_single.py tests/test_gpu_storage_borsh.py tests/test_gpu_borsh.py -x -q -m gpu > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02m_tests.log
python bench.py --workload config1 > gpurun_out/r02m_bench_config1.json 2> gpurun_out/r02m_bench_config1.err; echo "config1 rc=$?"; tail -3 gpurun_out/r02m_bench_config1.err
python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02m_bench_config1.json') if l.startswith('{')][0]); print({k:j[k] for k in ('latency_us','latency_us_c_abi','latency_us_borsh_c_abi','gpu_launches')})"
python bench.py --no-configs --no-cpu-baseline --steps 3 > gpurun_out/r02m_bench_quick.json 2> gpurun_out/r02m_bench_quick.err; echo "quick bench rc=$?"; tail -3 gpurun_out/r02m_bench_quick.err
python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02m_bench_quick.json') if l.startswith('{')][0]); e=j['e2e']; print(j['value'], e['value'], e['roofline']['bound'], e['roofline']['frac'])"
