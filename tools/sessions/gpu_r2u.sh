#!/bin/bash
# round 2, session u: full GPU suite on the reworked stand-alone checker, checker sweep, compute-sanitizer over every kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2u_pytest.log
python tools/r1cs_sweep.py 2>&1 | tee gpurun_out/r2u_r1cs_sweep.jsonl
rm -f gpurun_out/r2u_compute_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py" | tee -a gpurun_out/r2u_compute_sanitizer.txt
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done|Error|error|hazard" | head -20 | tee -a gpurun_out/r2u_compute_sanitizer.txt
done
