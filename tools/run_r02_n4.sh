#!/bin/bash
# N = 4: the driver's command line for the default bench under torchrun
cd /root/repo
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_scale_n4.err
