set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain4_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_config2.csv $CMD > gpurun_out/ncu4_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_keccak256_nodes|k_verify_fast" -s 6 -c 2 -o gpurun_out/r01d_config2 $CMD > gpurun_out/ncu4_k.log 2>&1
CMD4="python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD4 > gpurun_out/plain4_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches_config4.csv $CMD4 > gpurun_out/ncu4_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trie_structure|k_keccak256_leaves" -s 4 -c 2 -o gpurun_out/r01d_config4 $CMD4 > gpurun_out/ncu4_t.log 2>&1
ls -la gpurun_out | grep r01d
