#!/bin/bash
# end-of-round check of the final tree: full GPU suite, smoke, borsh stream sweep, config 2 with the borsh leg
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/final3_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final3_smoke.log
python tools/borsh_stream_bench.py > gpurun_out/borsh_sweep2.txt 2> gpurun_out/borsh_sweep2.err; echo "sweep rc=$?"; cat gpurun_out/borsh_sweep2.txt
python bench.py --borsh --steps 5 --warmup 3 > gpurun_out/r01_bench_config2.json 2> gpurun_out/f2.err; echo "bench rc=$?"
