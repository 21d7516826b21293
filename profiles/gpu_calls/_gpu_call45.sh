#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu45.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu45.log
tail -15 gpurun_out/pytest_gpu45.log
