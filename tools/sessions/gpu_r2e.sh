#!/bin/bash
# round 2: compute-sanitizer over every kernel (small batches), wide-path tests after the inverse-table change, fr throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nova_wide.py tests/test_gpu_fr_batches.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2e_pytest.log
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py" | tee -a gpurun_out/r2e_compute_sanitizer.txt
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done|Error|error|hazard" | head -20 | tee -a gpurun_out/r2e_compute_sanitizer.txt
done
