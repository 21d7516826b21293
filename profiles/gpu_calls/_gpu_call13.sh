#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu13.log
tail -4 gpurun_out/pytest_gpu13.log
python bench.py > gpurun_out/bench_r02f_n1.json 2> gpurun_out/bench_r02f_n1.err; echo "bench rc=$?"
head -c 400 gpurun_out/bench_r02f_n1.json; echo
python tests/_variant_bench_bin.py libdcsmc.so 16 10000
