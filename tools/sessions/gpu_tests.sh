#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python tools/nova_quickbench.py 2>&1 | tee gpurun_out/nova_quickbench.log
python tools/bench_configs.py 2>&1 | tee gpurun_out/configs.log
