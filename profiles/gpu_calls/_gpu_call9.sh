#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_leaf_propose2|k_seg_resample|k_expand|k_dart_partition|k_scan2' -c 12 -o gpurun_out/r02m_hot python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02m.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02m_metrics_cfg2.csv python bench.py --config cfg2 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02m_metrics.log 2>&1; echo "ncu metrics rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --sub none --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02_launches.log 2>&1; echo "ncu launches rc=$?"
wc -l gpurun_out/r02_launches_cfg3.csv gpurun_out/r02m_metrics_cfg2.csv
