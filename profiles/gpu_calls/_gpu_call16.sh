#!/bin/bash
mkdir -p gpurun_out
for v in 1 5 4; do
DCSMC_LEAFKIDS=$v timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose_lk' -c 1 -o gpurun_out/r02p_lk$v python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02p_$v.log 2>&1; echo "ncu lk$v rc=$?"
done
ls -la gpurun_out
