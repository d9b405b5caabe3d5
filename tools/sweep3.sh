#!/bin/bash
mkdir -p gpurun_out
for v in ncw6 ncw8; do
  echo "== $v"; B3W_EXP_LIB=$PWD/build_exp/lib_$v.so python tools/r1cs_quickbench.py 2>&1 | grep nova | tee gpurun_out/r1cs_quickbench_$v.log
done
