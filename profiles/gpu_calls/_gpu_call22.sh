#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose_lk' -c 1 -o gpurun_out/r02q_lk python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02q_lk.log 2>&1; echo "ncu lk rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose<' -c 1 -o gpurun_out/r02q_gen python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02q_gen.log 2>&1; echo "ncu gen rc=$?"
ls -la gpurun_out
