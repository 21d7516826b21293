#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kernel_variants" > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu15.log
tail -15 gpurun_out/pytest_gpu15.log
for v in 1 2 3 4 5; do DCSMC_LEAFKIDS=$v python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | sed "s/^/LK=$v /"; done | tee gpurun_out/lk_variants.log
for v in 1 3; do DCSMC_LEAFKIDS=$v python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | sed "s/^/LK=$v /"; done | tee -a gpurun_out/lk_variants.log
