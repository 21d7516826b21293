#!/bin/bash
mkdir -p gpurun_out
./profiles/microbench/issue_model | tee gpurun_out/issue_model.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu18.log
tail -5 gpurun_out/pytest_gpu18.log
python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1
python tests/_variant_bench.py libdcsmc.so 1 1000000 2>&1 | tail -1
