#!/bin/bash
mkdir -p gpurun_out
python tests/_repro_resend.py 2>&1 | tail -6 | tee gpurun_out/repro_resend.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu6.log
tail -8 gpurun_out/pytest_gpu6.log
python bench.py > gpurun_out/bench_r02c_n1.json 2> gpurun_out/bench_r02c_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r02c_n1.json
