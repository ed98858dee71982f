set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_config2_reference.json 2> gpurun_out/f_ref.err; tail -2 gpurun_out/f_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_config2.json 2> gpurun_out/f2.err; tail -2 gpurun_out/f2.err
python bench.py --workload config1 --steps 5 > gpurun_out/r01_bench_config1.json 2> gpurun_out/f1.err; tail -2 gpurun_out/f1.err
python bench.py --workload config3 --steps 3 > gpurun_out/r01_bench_config3.json 2> gpurun_out/f3.err; tail -2 gpurun_out/f3.err
python bench.py --workload config4 --steps 3 > gpurun_out/r01_bench_config4.json 2> gpurun_out/f4.err; tail -2 gpurun_out/f4.err
python bench.py --workload config5 --steps 2 > gpurun_out/r01_bench_config5.json 2> gpurun_out/f5.err; tail -2 gpurun_out/f5.err
python bench.py --steps 5 --dedup --no-e2e > gpurun_out/r01_bench_config2_dedup_secondary.json 2> gpurun_out/f2d.err; tail -2 gpurun_out/f2d.err
python -c "import __graft_entry__ as g; g.smoke()"
