#!/bin/bash
# round 2, first GPU session: the new tests first (fail fast), then the whole suite, then the checker timings
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2a_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_compressible.py tests/test_gpu_extras.py tests/test_gpu_fr_batches.py tests/test_gpu_r1cs.py tests/test_gpu_chain.py tests/test_gpu_hybrid.py tests/test_gpu_nova_wide.py -q -m gpu -x 2>&1 | tail -40 | tee gpurun_out/r2a_pytest_new.log
timeout 300 python tools/r1cs_quickbench.py 2>&1 | tee gpurun_out/r2a_r1cs.jsonl
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2a_pytest_all.log
