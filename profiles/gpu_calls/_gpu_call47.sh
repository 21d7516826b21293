#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu47.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu47.log
tail -6 gpurun_out/pytest_gpu47.log
python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | tee gpurun_out/variants47.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/variants47.log
python tests/_variant_bench_bin.py libdcsmc.so 16 10000 | tee -a gpurun_out/variants47.log
