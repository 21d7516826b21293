#!/bin/bash
mkdir -p gpurun_out
for v in 2 6 7 8 9; do
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 100000
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 1000000
done 2>&1 | tee gpurun_out/scan_variants2.log
DCSMC_HOSTPROF=1 python tests/_e2e_profile_bin.py 16 10000 2>&1 | tail -60 | tee gpurun_out/e2e_bin_profile.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scan2|k_seg_resample|k_dart_partition|k_expand" -c 8 \
  -o gpurun_out/r02b_chain python tests/_variant_bench.py libdcsmc.so 0 1000000 > gpurun_out/ncu_chain.log 2>&1; echo "ncu rc=$?"
