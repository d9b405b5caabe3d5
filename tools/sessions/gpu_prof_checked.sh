#!/bin/bash
# GPU session: ncu --set full captures of the nova kernel and of the fused-check kernels (one launch each), plus the
# quick device timings of all variants.
mkdir -p gpurun_out
python tools/nova_quickbench.py 2>&1 | tee gpurun_out/nova_quickbench.log
python tools/r1cs_quickbench.py 2>&1 | tee gpurun_out/r1cs_quickbench.log
for spec in "blake3_nova plain" "blake3_nova checked" "blake3_compression checked"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blake3 -s 2 -c 1 -f \
    -o gpurun_out/prof_${1}_${2} python tools/prof_run.py 15 4 $1 $2 > gpurun_out/ncu_${1}_${2}.log 2>&1
  tail -2 gpurun_out/ncu_${1}_${2}.log
done
ls -la gpurun_out
