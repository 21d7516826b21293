#!/bin/bash
mkdir -p gpurun_out
{
for lib in libdcsmc.so libdcsmc_x.so libdcsmc_y.so; do
  python tests/_variant_bench.py $lib 0 100000
  python tests/_variant_bench.py $lib 0 1000000
  python tests/_variant_bench_bin.py $lib 14 10000
done
} 2>&1 | tee gpurun_out/fold_ab.log
