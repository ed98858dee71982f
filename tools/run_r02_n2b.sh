#!/bin/bash
# N = 2 on the tree with the device-flatten e2e leg: multi-GPU tests, then the default bench line under torchrun
cd /root/repo
nproc
python -m pytest tests/test_gpu_multi.py tests/test_gpu_borsh.py tests/test_gpu_rebuild.py -q -m gpu > gpurun_out/r02b_n2_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r02b_n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02b_scale_n2.json 2> gpurun_out/r02b_scale_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b_scale_n2.err
