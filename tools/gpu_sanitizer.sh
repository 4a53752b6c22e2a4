#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dedisperse" 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit $?"; tail -4 gpurun_out/sanitizer_$tool.txt
done
