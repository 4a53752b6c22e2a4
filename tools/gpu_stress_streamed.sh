#!/bin/bash
# stress the streamed-weights kernel at the shapes that exposed the shared-ring race, then parity + timing
mkdir -p gpurun_out
export N_LAUNCH=200
for s in "256 2048 256 8" "256 2048 256 16" "256 1024 256 16" "256 1536 256 8" "192 2048 256 16"; do
  echo "== stress $s"; timeout 200 python tools/stress_launches.py $s 2>&1 | grep -v Warning | grep launch | tail -1 | cut -c1-200
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
DSABF_FORCE_STREAMED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for s in "256 2048 256 32" "128 1024 256 32" "64 512 256 32" "64 256 256 32"; do
  timeout 120 python tools/time_shape.py $s 2>&1 | tail -1
done | tee gpurun_out/time_shapes.txt
