set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01_scale_n${N}_config2.json 2> gpurun_out/n${N}.err; tail -2 gpurun_out/n${N}.err; head -c 250 gpurun_out/r01_scale_n${N}_config2.json
if [ "$2" = "c5" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 3 --warmup 3 > gpurun_out/r01_scale_n${N}_config5.json 2> gpurun_out/n${N}_c5.err; tail -2 gpurun_out/n${N}_c5.err; head -c 250 gpurun_out/r01_scale_n${N}_config5.json
fi
