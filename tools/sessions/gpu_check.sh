#!/bin/bash
# Quick GPU session: parity tests + both bench arms.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err; tail -3 gpurun_out/bench_own.err; cat gpurun_out/bench_own.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
