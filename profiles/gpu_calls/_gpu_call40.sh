#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_node_fused' -s 2 -c 1 -o gpurun_out/r02w_fused python tests/_variant_bench_bin.py libdcsmc.so 14 10000 > gpurun_out/ncu_r02w_fused.log 2>&1; echo "ncu fused rc=$?"
tail -2 gpurun_out/ncu_r02w_fused.log
