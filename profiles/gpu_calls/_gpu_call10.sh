#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dc_api.py -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu10.log
tail -4 gpurun_out/pytest_gpu10.log
{
DCSMC_LEAF_V=2 python tests/_variant_bench.py libdcsmc.so 0 100000
for ch in 1024 2048 4096; do
  DCSMC_LEAF_V=3 DCSMC_LEAF_CH=$ch python tests/_variant_bench.py libdcsmc.so 0 100000
done
DCSMC_LEAF_V=2 python tests/_variant_bench.py libdcsmc.so 0 1000000
for ch in 2048 4096; do
  DCSMC_LEAF_V=3 DCSMC_LEAF_CH=$ch python tests/_variant_bench.py libdcsmc.so 0 1000000
done
DCSMC_LEAF_V=2 python tests/_variant_bench.py libdcsmc.so 1 100000
DCSMC_LEAF_V=3 python tests/_variant_bench.py libdcsmc.so 1 100000
} 2>&1 | tee gpurun_out/leaf3_ab.log
