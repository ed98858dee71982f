#!/bin/bash
# the round's last session: whole GPU suite, smoke, the default bench line, on the final tree
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r02z_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02z_smoke.log
python bench.py > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02z_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; echo "reference rc=$?"
