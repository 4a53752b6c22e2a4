#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or scaled" 2>&1 | tail -2
DSABF_FORCE_STREAMED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for s in "256 2048 256 32" "128 1024 256 32" "64 512 256 32" "128 256 256 32" "192 768 256 32"; do
  timeout 120 python tools/time_shape.py $s 2>&1 | tail -1 | cut -c1-190
done
N_LAUNCH=100 timeout 200 python tools/stress_launches.py 128 2048 256 16 2>&1 | grep launch | tail -1
N_LAUNCH=100 timeout 200 python tools/stress_launches.py 64 1024 256 32 2>&1 | grep launch | tail -1
