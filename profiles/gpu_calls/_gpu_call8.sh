#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu8.log
tail -4 gpurun_out/pytest_gpu8.log
for lib in libdcsmc.so libdcsmc_dk.so libdcsmc_l5.so; do
  for v in 1 2; do
    DCSMC_LEAF_V=$v python tests/_variant_bench.py $lib 0 100000
  done
done 2>&1 | tee gpurun_out/leaf_variants8.log
python bench.py > gpurun_out/bench_r02d_n1.json 2> gpurun_out/bench_r02d_n1.err; echo "bench rc=$?"
head -c 1500 gpurun_out/bench_r02d_n1.json
