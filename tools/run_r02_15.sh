#!/bin/bash
cd /root/repo
for w in bare torch omp1 csr bare; do python tools/e2e_context_probe.py $w 2>/dev/null | tail -1; done
python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02e_bench.json 2>gpurun_out/r02e_bench.err; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02e_bench.json') if l.startswith('{')][0]); e=j['e2e']; print('bench e2e', e['ms_per_step'], e['step_ms_rank0'], e['host_ms'])"
OMP_NUM_THREADS=1 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r02e_bench_omp1.json 2>gpurun_out/r02e_bench_omp1.err; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02e_bench_omp1.json') if l.startswith('{')][0]); e=j['e2e']; print('bench e2e OMP=1', e['ms_per_step'], e['step_ms_rank0'], e['host_ms'])"
