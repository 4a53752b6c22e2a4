#!/bin/bash
mkdir -p gpurun_out
for s in "256 2048 256 8" "256 2048 256 32" "128 1024 256 32" "64 512 256 32" "256 256 256 32" "64 256 256 32"; do
  timeout 120 python tools/time_shape.py $s 2>&1 | tail -1
done | tee gpurun_out/time_shapes.txt
echo forced; DSABF_FORCE_STREAMED=1 timeout 120 python tools/time_shape.py 64 256 256 32 | tee -a gpurun_out/time_shapes.txt
