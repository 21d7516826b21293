#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_gpu41.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu41.log
tail -3 gpurun_out/pytest_gpu41.log
for pf in 3 0; do DCSMC_L2_PREFETCH=$pf python tests/_variant_bench_bin.py libdcsmc.so 16 10000 | sed "s/^/PF=$pf /"; done | tee gpurun_out/variants41.log
