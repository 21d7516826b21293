#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu26.log
tail -3 gpurun_out/pytest_gpu26.log
for lib in libdcsmc.so libdcsmc_i5.so; do
python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1
DCSMC_LEAFKIDS=0 python tests/_variant_bench.py $lib 0 100000 2>&1 | tail -1 | sed "s/^/LK=0 /"
done | tee gpurun_out/variants26.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/variants26.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose' -s 1 -c 1 -o gpurun_out/r02r_gen python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02r_gen.log 2>&1; echo "ncu gen rc=$?"
