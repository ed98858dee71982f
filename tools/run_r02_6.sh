#!/bin/bash
# round 2, session 6: K2f v2 + K2b with shared-memory columns: parity, then the walk on config 2 / shuffled config 2 / config 3
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_single.py tests/test_gpu_storage.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r02_6_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_6_tests.log
for lib in zk-state-proofs_b200/libmptv.so build/variants/libmptv_walk3.so build/variants/libmptv_walk2.so; do
  for w in "--workload config2" "--workload config2 --shuffled" "--workload config3"; do
    MPTV_LIB=$lib python bench.py $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', 'ms/step %.3f' % d['ms_per_step'], {k: round(v,3) for k,v in d['kernel_ms'].items()}, d.get('verdicts'))"
  done
done
MPTV_LIB=build/variants/libmptv_smalltiming.so python tools/latency_probe.py
