#!/bin/bash
cd /root/repo
python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 --l2-fetch 32 > gpurun_out/l2f_plain.json 2> gpurun_out/l2f_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_keccak256_nodes --launch-skip 1 -c 1 -o gpurun_out/l2f32_k1 -f \
  python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 --l2-fetch 32 > gpurun_out/l2f_ncu.log 2>&1
echo done
