set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01_scale_n2_config2.json 2> gpurun_out/n2.err; tail -5 gpurun_out/n2.err; cat gpurun_out/r01_scale_n2_config2.json | head -c 600
