#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for spec in "blake3_nova checked" "blake3_compression checked"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3 -s 2 -c 1 -f \
    -o gpurun_out/prof2_${1}_${2} python tools/prof_run.py 15 4 $1 $2 > gpurun_out/ncu2_${1}_${2}.log 2>&1
  tail -1 gpurun_out/ncu2_${1}_${2}.log
done
python tools/bench_configs.py 3 4 2>&1 | tee gpurun_out/configs34.log
