#!/bin/bash
# scratch driver of one gpurun call: parity tests, scan variants A/B, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in 1 2 3 4 5; do
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 100000
  DCSMC_SCAN_V=$v python tests/_variant_bench.py libdcsmc.so 0 1000000
done 2>&1 | tee gpurun_out/scan_variants.log
python bench.py > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02b_n1.json
