#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_gpu35.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu35.log
tail -3 gpurun_out/pytest_gpu35.log
for v in 1 2; do
DCSMC_LEAFKIDS=$v python tests/_variant_bench.py libdcsmc_lk4.so 0 100000 2>&1 | tail -1 | sed "s/^/LK=$v /"
done | tee gpurun_out/lk4_variants.log
python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | tee -a gpurun_out/lk4_variants.log
DCSMC_LEAFKIDS=2 python tests/_variant_bench.py libdcsmc_lk4.so 0 1000000 2>&1 | tail -1 | sed "s/^/LK=2 /" | tee -a gpurun_out/lk4_variants.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/lk4_variants.log
