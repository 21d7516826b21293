#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/pytest_gpu48.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu48.log
tail -3 gpurun_out/pytest_gpu48.log
python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | tee gpurun_out/variants48.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/variants48.log
