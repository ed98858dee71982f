#!/bin/bash
# 2 GPUs, final tree: whole GPU suite (the multi-GPU tests run here), then the driver's N = 2 command
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r02f_gpu_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02f_gpu_tests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02f_scale_n2.json 2> gpurun_out/r02f_scale_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_scale_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_scale_n2_reference.json 2> gpurun_out/r02f_scale_n2_reference.err; echo "reference rc=$?"; cat gpurun_out/r02f_scale_n2_reference.json | cut -c1-300
