#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_packed.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python tools/packed_quickbench.py 2>&1 | tee gpurun_out/packed_quickbench.log
