#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform -s 2 -c 1 -f -o gpurun_out/prof_obs4 python tools/prof_one.py obs 0 4 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform -s 2 -c 1 -f -o gpurun_out/prof_dbg4 python tools/prof_one.py dbg 0 4 2>&1 | tail -1
