#!/bin/bash
mkdir -p gpurun_out
./profiles/microbench/issue_model | tee gpurun_out/issue_model.log
for lib in libdcsmc.so libdcsmc_z1.so libdcsmc_z2.so libdcsmc_z1c.so libdcsmc_c5.so; do
python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1
done | tee gpurun_out/expand_variants.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu19.log
tail -3 gpurun_out/pytest_gpu19.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/expand_variants.log
python tests/_variant_bench_bin.py libdcsmc.so 16 10000 | tee -a gpurun_out/expand_variants.log
