#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_dc_api.py -m gpu -q -x -k "sharded" > gpurun_out/pytest_gpu31.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu31.log
tail -4 gpurun_out/pytest_gpu31.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r02j_n4.json 2> gpurun_out/bench_r02j_n4.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r02j_n4.json | head -c 400; echo
