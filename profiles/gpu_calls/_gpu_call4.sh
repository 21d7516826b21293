#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --report-api-errors all --print-limit 20 python -m pytest tests/test_gpu_dc_api.py -x -q -k "resending_the_same_tree" > gpurun_out/sanitizer_resend.log 2>&1
grep -n "Program hit\|=========     at\|Host Frame.*dcsmc\|invalid" gpurun_out/sanitizer_resend.log | head -40
for lk in 0 1; do
  DCSMC_LEAFKIDS=$lk python tests/_variant_bench.py libdcsmc.so 0 100000
  DCSMC_LEAFKIDS=$lk python tests/_variant_bench.py libdcsmc.so 0 1000000
done 2>&1 | tee gpurun_out/leafkids.log
DCSMC_HOSTPROF=1 python tests/_e2e_profile_bin.py 16 10000 2>&1 | tail -4 | tee gpurun_out/e2e_bin_profile3.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu4.log
tail -8 gpurun_out/pytest_gpu4.log
