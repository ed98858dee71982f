set -x
CMD="python bench.py --accounts 2000000 --proofs 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_config2.csv $CMD > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_keccak256_nodes -s 3 -c 2 -o gpurun_out/r01b_keccak $CMD > gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_verify_walk -s 3 -c 2 -o gpurun_out/r01b_walk $CMD > gpurun_out/ncu_w.log 2>&1
CMD4="python bench.py --workload config4 --blocks 1000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches_config4.csv $CMD4 > gpurun_out/ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trie_ -s 9 -c 6 -o gpurun_out/r01b_trie $CMD4 > gpurun_out/ncu_t.log 2>&1
ls -la gpurun_out
