#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu21.log
tail -3 gpurun_out/pytest_gpu21.log
for pf in 1 0; do DCSMC_L2_PREFETCH=$pf python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | sed "s/^/PF=$pf /"; done | tee gpurun_out/variants21.log
DCSMC_LEAFKIDS=0 python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | sed "s/^/LK=0 /" | tee -a gpurun_out/variants21.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/variants21.log
python tests/_variant_bench_bin.py libdcsmc.so 16 10000 | tee -a gpurun_out/variants21.log
