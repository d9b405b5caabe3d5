#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list of the bench command and one full capture of the witness kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err; tail -3 gpurun_out/bench_own.err; cat gpurun_out/bench_own.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_blake3 -s 2 -c 1 -f -o gpurun_out/prof_comp python tools/prof_run.py 16 4 > gpurun_out/ncu_full.log 2>&1; tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
