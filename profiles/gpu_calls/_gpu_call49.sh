#!/bin/bash
# final state: whole GPU suite, default bench line, smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu49.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu49.log
tail -4 gpurun_out/pytest_gpu49.log
python bench.py > gpurun_out/bench_r02p_n1.json 2> gpurun_out/bench_r02p_n1.err; echo "bench rc=$?"
head -c 300 gpurun_out/bench_r02p_n1.json; echo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
