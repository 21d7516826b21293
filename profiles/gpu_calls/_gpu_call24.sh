#!/bin/bash
mkdir -p gpurun_out
for lib in libdcsmc.so libdcsmc_l8.so libdcsmc_l9.so libdcsmc_l10.so libdcsmc.so libdcsmc_l9.so; do
python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1
done | tee gpurun_out/leaf_shape.log
