#!/bin/bash
mkdir -p gpurun_out
for lib in libdcsmc.so libdcsmc_gs.so; do
python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1
DCSMC_LEAFKIDS=0 python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1 | sed "s/^/LK=0 /"
python tests/_variant_bench.py $lib 0 1000000 2>&1 | tail -1
done | tee gpurun_out/gs_variants.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/pytest_gpu34.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu34.log
tail -3 gpurun_out/pytest_gpu34.log
