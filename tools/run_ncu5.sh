set -x
CMD4="python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD4 > gpurun_out/plain5_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01e_launches_config4.csv $CMD4 > gpurun_out/ncu5_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trie_structure|k_keccak256_leaves|k_trie_encode" -s 8 -c 4 -o gpurun_out/r01e_config4 $CMD4 > gpurun_out/ncu5_t.log 2>&1
