#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4 5 6 8; do
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 100000
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 1000000
done 2>&1 | tee gpurun_out/scan_variants3.log
DCSMC_HOSTPROF=1 python tests/_e2e_profile_bin.py 16 10000 2>&1 | tail -12 | tee gpurun_out/e2e_bin_profile2.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu3.log
tail -5 gpurun_out/pytest_gpu3.log
