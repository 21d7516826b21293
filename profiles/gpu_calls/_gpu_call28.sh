#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu28.log
tail -4 gpurun_out/pytest_gpu28.log
python bench.py > gpurun_out/bench_r02i_n1.json 2> gpurun_out/bench_r02i_n1.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench_r02i_n1.json; echo
