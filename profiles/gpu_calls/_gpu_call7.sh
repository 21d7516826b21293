#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu7.log
tail -4 gpurun_out/pytest_gpu7.log
for lib in libdcsmc_prev.so libdcsmc.so; do
  for v in 1 2; do
    DCSMC_LEAF_V=$v python tests/_variant_bench.py $lib 0 100000
  done
  python tests/_variant_bench.py $lib 1 100000
  python tests/_variant_bench_bin.py $lib 14 10000
done 2>&1 | tee gpurun_out/divsafe_ab.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_leaf_propose2|k_internal_propose' -c 4 -o gpurun_out/r02k_leaf_merge python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02k.log 2>&1; echo "ncu rc=$?"
