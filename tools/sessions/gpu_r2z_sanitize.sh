#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over every kernel, final code of round 2
mkdir -p gpurun_out
rm -f gpurun_out/r2z_compute_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py" | tee -a gpurun_out/r2z_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done|Error|error|hazard" | head -20 | tee -a gpurun_out/r2z_compute_sanitizer.txt
done
