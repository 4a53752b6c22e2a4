#!/bin/bash
mkdir -p gpurun_out
for shape in obs dbg; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform -s 2 -c 1 -f -o gpurun_out/prof_$shape python tools/prof_one.py $shape 0 4 2>&1 | tail -3
done
ls -la gpurun_out/*.ncu-rep
