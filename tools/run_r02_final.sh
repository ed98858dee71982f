#!/bin/bash
# final tree: whole GPU suite, smoke, the default bench line (driver's N = 1 command), the reference arm, K1 re-capture
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r02c_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02c_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02c_smoke.log
python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02c_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err; echo "reference rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$B > gpurun_out/r02c_ncu_k1_plain.json 2> gpurun_out/r02c_ncu_k1_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_keccak256_nodes --launch-skip 1 -c 1 -o gpurun_out/r02c_ncu_k1 $B > gpurun_out/r02c_ncu_k1.log 2>&1; echo "k1 ncu rc=$?"
ncu -i gpurun_out/r02c_ncu_k1.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_k1_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02c | awk '{print $5, $9}'
