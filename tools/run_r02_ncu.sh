#!/bin/bash
# round 2 ncu evidence: launch list of the default command (device-resident part), then --set full captures of
# K1, K2f (config 2), K2b (config 2 shuffled = 100 % deferred, and config 3).  Each capture follows a plain run of the same command.
cd /root/repo
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
run_cap () {  # name, kernel regex, extra bench args
  $B $3 > gpurun_out/r02_ncu_$1_plain.json 2> gpurun_out/r02_ncu_$1_plain.err || { echo "$1: plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -o gpurun_out/r02_ncu_$1 $B $3 > gpurun_out/r02_ncu_$1.log 2>&1
  echo "$1 ncu rc=$?"
  ncu -i gpurun_out/r02_ncu_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
}
$B > gpurun_out/r02_launches_plain.json 2> gpurun_out/r02_launches_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
run_cap k1 k_keccak256_nodes ""
run_cap k2f k_verify_fast ""
run_cap k2b_shuffled k_verify_walk "--shuffled"
run_cap k2b_config3 k_verify_walk "--workload config3"
ls -la gpurun_out | grep r02_ncu | awk '{print $5, $9}'
