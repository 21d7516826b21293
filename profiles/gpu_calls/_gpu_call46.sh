#!/bin/bash
mkdir -p gpurun_out
python tests/_div_probe.py 2>&1 | tee gpurun_out/div_probe.log
