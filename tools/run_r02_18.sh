#!/bin/bash
# final tree: whole GPU suite, smoke, default bench, config 3 full line, storage stream campaign
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r02j_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02j_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02j_smoke.log
python bench.py > gpurun_out/r02j_bench_default.json 2> gpurun_out/r02j_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02j_bench_default.err
python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/r02j_bench_config3.json 2> gpurun_out/r02j_bench_config3.err; echo "config3 rc=$?"
timeout 200 python tools/fuzz_storage_stream.py 150 > gpurun_out/r02j_fuzz_storage_stream.txt 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r02j_fuzz_storage_stream.txt
