#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose_lk' -c 1 -o gpurun_out/r02t_lk python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02t_lk.log 2>&1; echo "ncu lk rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_internal_propose' -s 1 -c 1 -o gpurun_out/r02t_gen python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02t_gen.log 2>&1; echo "ncu gen rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02t_metrics_cfg2.csv python bench.py --config cfg2 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02t_metrics.log 2>&1; echo "ncu metrics rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02t_launches_cfg3.csv python bench.py --sub none --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02t_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out | tail -8
