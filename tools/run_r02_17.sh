#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_storage_borsh.py -x -q -m gpu > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02h_tests.log
python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/r02h_bench_config3.json 2> gpurun_out/r02h_bench_config3.err; echo "bench rc=$?"; tail -5 gpurun_out/r02h_bench_config3.err
python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02h_bench_config3.json') if l.startswith('{')][0]); e=j['e2e']; print('config3', j['value'], j['kernel_ms'], j['verdicts']); print('e2e', e['entry'], e['value'], e['ms_per_step'], e['step_ms_rank0'], e['h2d_bytes_per_step'], e['input_bytes_per_step'], e['host_ms'], e['transfer_dedup']); print('csr', j['e2e_csr']['value'], j['e2e_csr']['ms_per_step'])"
