set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --accounts 1000000 --proofs 100000 --steps 3 --warmup 3 > gpurun_out/b2_small.json 2> gpurun_out/b2_small.err; tail -3 gpurun_out/b2_small.err
python bench.py --workload config3 --accounts 1000000 --proofs 200000 --slots 100000 --tokens 4 --steps 3 > gpurun_out/b3_small.json 2> gpurun_out/b3_small.err; tail -3 gpurun_out/b3_small.err
python bench.py --workload config5 --accounts 1000000 --proofs 200000 --total-proofs 800000 --slots 100000 --tokens 4 --steps 3 > gpurun_out/b5_small.json 2> gpurun_out/b5_small.err; tail -3 gpurun_out/b5_small.err
python bench.py --workload config4 --blocks 500 --steps 3 > gpurun_out/b4_small.json 2> gpurun_out/b4_small.err; tail -3 gpurun_out/b4_small.err
nproc; free -g | head -2
