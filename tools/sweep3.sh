#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_own2.json 2> gpurun_out/bench_own2.err; tail -3 gpurun_out/bench_own2.err; cat gpurun_out/bench_own2.json
