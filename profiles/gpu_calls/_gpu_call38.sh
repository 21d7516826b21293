#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r02m_n8.json 2> gpurun_out/bench_r02m_n8.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r02m_n8.json | head -c 400; echo
tail -3 gpurun_out/bench_r02m_n8.err
