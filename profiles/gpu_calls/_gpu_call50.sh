#!/bin/bash
mkdir -p gpurun_out
for ch in 4096 2048 8192 16384; do DCSMC_LEAF_CH=$ch python tests/_variant_bench.py libdcsmc.so 0 1000000 2>&1 | tail -1 | sed "s/^/CH=$ch /"; done | tee gpurun_out/leaf_ch.log
