#!/bin/bash
mkdir -p gpurun_out
for lib in libdcsmc.so libdcsmc_lk4.so libdcsmc_lk6.so libdcsmc_i4.so libdcsmc.so; do
python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1
done | tee gpurun_out/minb_variants.log
for lib in libdcsmc.so libdcsmc_lk4.so libdcsmc_i4.so; do
python tests/_variant_bench.py $lib 0 1000000 2>&1 | tail -1
done | tee -a gpurun_out/minb_variants.log
