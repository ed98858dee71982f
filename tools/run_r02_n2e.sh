#!/bin/bash
cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-configs > gpurun_out/r02n_scale_n2_quick.json 2> gpurun_out/r02n_scale_n2_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r02n_scale_n2_quick.err
python -c "
import json; j=json.loads([l for l in open('gpurun_out/r02n_scale_n2_quick.json') if l.startswith('{')][0]); e=j['e2e']; print(j['value'], e['value'], {k:e['roofline'][k] for k in ('bound','achieved','peak','frac')})"
