#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or scaled" 2>&1 | tail -3
DSABF_FORCE_STREAMED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for s in "256 2048 256 32" "256 2048 256 8" "128 1024 256 32" "64 512 256 32" "256 256 256 32"; do
  timeout 120 python tools/time_shape.py $s 2>&1 | tail -1
done | tee gpurun_out/time_shapes.txt
N_LAUNCH=100 timeout 200 python tools/dbg_launches.py 256 2048 256 16 2>&1 | grep launch | tail -1
