#!/bin/bash
# N = 2: the default bench line under torchrun (weak-scaled config 2 + single_context + config5), then the multi-GPU tests
cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_scale_n2.err
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_n2_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r02_n2_tests.log
