set -x
# full-size config 2: launch list + full capture of K1 (traffic per launch at the bench's own size) and the two walk kernels
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain2_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_config2.csv $CMD > gpurun_out/ncu2_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_keccak256_nodes|k_verify_fast|k_verify_walk" -s 9 -c 3 -o gpurun_out/r01c_config2 $CMD > gpurun_out/ncu2_k.log 2>&1
# config 4 at 2000 blocks: launch list + full capture of the structure kernel, the fused leaf hash and the encode of level 1
CMD4="python bench.py --workload config4 --blocks 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD4 > gpurun_out/plain2_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches_config4.csv $CMD4 > gpurun_out/ncu2_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trie_structure|k_keccak256_leaves|k_trie_encode" -s 6 -c 3 -o gpurun_out/r01c_config4 $CMD4 > gpurun_out/ncu2_t.log 2>&1
ls -la gpurun_out | tail -12
