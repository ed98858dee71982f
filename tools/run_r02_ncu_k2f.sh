#!/bin/bash
# ncu --set full capture of K2f (k_verify_fast) on config 2, after a plain run of the same command
cd /root/repo
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > gpurun_out/r02_k2f_plain.json 2> gpurun_out/r02_k2f_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r02_k2f_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_verify_fast -c 1 -o gpurun_out/r02_k2f $CMD > gpurun_out/r02_k2f_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_k2f.ncu-rep --page raw --csv > gpurun_out/r02_k2f_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep k2f
