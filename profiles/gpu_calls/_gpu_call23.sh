#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu23.log
tail -3 gpurun_out/pytest_gpu23.log
for pf in 3 2 1 0; do DCSMC_L2_PREFETCH=$pf python tests/_variant_bench.py libdcsmc.so 0 100000 2>&1 | tail -1 | sed "s/^/PF=$pf /"; done | tee gpurun_out/variants23.log
python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | tee -a gpurun_out/variants23.log
DCSMC_L2_PREFETCH=0 python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | sed "s/^/PF=0 /" | tee -a gpurun_out/variants23.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_leaf_propose3' -c 1 -o gpurun_out/r02q_leaf3 python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02q_leaf3.log 2>&1; echo "ncu leaf rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose' -s 1 -c 1 -o gpurun_out/r02q_gen python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02q_gen.log 2>&1; echo "ncu gen rc=$?"
