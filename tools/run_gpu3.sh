set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_config2.json 2> gpurun_out/b2.err; tail -3 gpurun_out/b2.err
python bench.py --workload config4 --steps 3 > gpurun_out/r01_bench_config4.json 2> gpurun_out/b4.err; tail -3 gpurun_out/b4.err
python bench.py --workload config3 --steps 3 > gpurun_out/r01_bench_config3.json 2> gpurun_out/b3.err; tail -3 gpurun_out/b3.err
python bench.py --workload config5 --steps 2 > gpurun_out/r01_bench_config5.json 2> gpurun_out/b5.err; tail -3 gpurun_out/b5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_config2_reference.json 2> gpurun_out/b2r.err; tail -3 gpurun_out/b2r.err
