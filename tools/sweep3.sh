#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_chain.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python tools/chain_quickbench.py 8192 2>&1 | tee gpurun_out/chain_quickbench2.log
