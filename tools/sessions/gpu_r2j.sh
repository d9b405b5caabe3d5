#!/bin/bash
# round 2, session j: fresh ncu captures of the stand-alone checker (current build: 4 loads in flight) for the two nova systems
mkdir -p gpurun_out
for c in blake3_nova_pasta blake3_nova_o1 blake3_compression; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_check_fast -s 1 -c 1 -f -o gpurun_out/r2j_prof_r1cs_$c python tools/prof_run.py 15 3 $c r1cs > gpurun_out/r2j_ncu_$c.log 2>&1; tail -2 gpurun_out/r2j_ncu_$c.log
done
ls -la gpurun_out | tail -5
