#!/bin/bash
mkdir -p gpurun_out
python tests/_repro_resend.py 2>&1 | tee gpurun_out/repro_resend.log
for v in 1 2; do
  DCSMC_LEAF_V=$v python tests/_variant_bench.py libdcsmc.so 0 100000
  DCSMC_LEAF_V=$v python tests/_variant_bench.py libdcsmc.so 0 1000000
  DCSMC_LEAF_V=$v python tests/_variant_bench.py libdcsmc.so 1 100000
done 2>&1 | tee gpurun_out/leafv.log
DCSMC_HOSTPROF=1 python tests/_e2e_profile_bin.py 16 10000 2>&1 | tail -4 | tee gpurun_out/e2e_bin_profile4.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu5.log
tail -8 gpurun_out/pytest_gpu5.log
