#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python tools/stream_sweep.py blake3_compression 16 checked 2>&1 | tee gpurun_out/sweep7_comp_chk.log
python tools/stream_sweep.py blake3_nova 16 checked 2>&1 | tee gpurun_out/sweep7_nova_chk.log
