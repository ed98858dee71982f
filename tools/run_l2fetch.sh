#!/bin/bash
# L2 fetch granularity hint: bench phases at 32 / 64 / 128 B, then one full ncu capture of K1 at 32 B
cd /root/repo
for g in 32 64 128; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 --l2-fetch $g > gpurun_out/l2f_config2_$g.json 2> gpurun_out/l2f_config2_$g.err || exit 1
done
for g in 32 128; do
  python bench.py --workload config4 --no-cpu-baseline --no-e2e --steps 3 --warmup 3 --l2-fetch $g > gpurun_out/l2f_config4_$g.json 2> gpurun_out/l2f_config4_$g.err || exit 1
done
python bench.py --workload config3 --no-cpu-baseline --no-e2e --steps 3 --warmup 3 --l2-fetch 32 > gpurun_out/l2f_config3_32.json 2> gpurun_out/l2f_config3_32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_keccak256_nodes -c 1 -o gpurun_out/l2f32_k1 -f \
  python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 --l2-fetch 32 > gpurun_out/l2f_ncu.log 2>&1
echo done
