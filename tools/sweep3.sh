#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_r1cs.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python tools/r1cs_quickbench.py 2>&1 | tee gpurun_out/r1cs_quickbench5.log
