#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu32.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu32.log
tail -4 gpurun_out/pytest_gpu32.log
python bench.py > gpurun_out/bench_r02k_n1.json 2> gpurun_out/bench_r02k_n1.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench_r02k_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02k_ref.json 2> gpurun_out/bench_r02k_ref.err; echo "ref rc=$?"
head -c 300 gpurun_out/bench_r02k_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02u_launches_cfg3.csv python bench.py --sub none --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r02u_launches.log 2>&1; echo "ncu launches rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
