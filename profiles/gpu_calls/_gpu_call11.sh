#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu11.log
tail -4 gpurun_out/pytest_gpu11.log
python bench.py > gpurun_out/bench_r02e_n1.json 2> gpurun_out/bench_r02e_n1.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_r02e_n1.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_leaf_propose3' -c 1 -o gpurun_out/r02n_leaf3 python tests/_variant_bench.py libdcsmc.so 0 100000 > gpurun_out/ncu_r02n.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02n_metrics_cfg2.csv python bench.py --config cfg2 --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02n_metrics.log 2>&1; echo "ncu metrics rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --sub none --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02_launches.log 2>&1; echo "ncu launches rc=$?"
