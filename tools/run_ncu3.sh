set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --dedup"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01d_launches_dedup.csv $CMD > gpurun_out/ncu3_l.log 2>&1
