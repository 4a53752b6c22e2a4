#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit $?"; tail -6 gpurun_out/sanitizer_$tool.txt
done
