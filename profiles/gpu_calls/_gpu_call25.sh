#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_dc_api.py -m gpu -q -x -k "sharded" > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu25.log
tail -4 gpurun_out/pytest_gpu25.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02h_n2.json 2> gpurun_out/bench_r02h_n2.err; echo "bench rc=$?"
head -c 700 gpurun_out/bench_r02h_n2.json; echo
