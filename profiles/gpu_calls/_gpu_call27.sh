#!/bin/bash
mkdir -p gpurun_out
for k in k_scan2 k_dart_partition k_seg_resample k_expand; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/r02s_$k python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02s_$k.log 2>&1; echo "ncu $k rc=$?"
done
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee gpurun_out/variants27.log
