#!/bin/bash
# N = 8: the driver's command line for the default bench under torchrun
cd /root/repo
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_scale_n8.err
