#!/bin/bash
# the driver's own invocation of the bench (round 1 used --steps 20 --warmup 5)
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02o_n1_s20.json 2> gpurun_out/bench_r02o_n1_s20.err ) 2>&1 | tail -3; echo "bench rc=$?"
head -c 300 gpurun_out/bench_r02o_n1_s20.json; echo
