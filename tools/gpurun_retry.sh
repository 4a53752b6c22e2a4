#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- <command>   : retries while the pod answers "transient"/busy
t=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --timeout $t "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
