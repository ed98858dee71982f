#!/bin/bash
# end-of-round check of the final code state: GPU suite, smoke, both bench arms, the dedup secondary figure
cd /root/repo
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/final2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final2_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_config2_reference.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_config2.json 2> gpurun_out/f2.err; echo "bench rc=$?"
python bench.py --steps 5 --dedup --no-e2e --no-cpu-baseline > gpurun_out/r01_bench_config2_dedup_secondary.json 2> gpurun_out/f2d.err; echo "dedup rc=$?"
