#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_obs_final python tools/prof_one.py obs 0 4 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_big -s 2 -c 1 -f -o gpurun_out/r2_prof_big_final python tools/prof_one.py big 0 4 2>&1 | tail -1
DD_T=2048 DD_B=32 DD_NDM=1627 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dedisperse_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_dd1627 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-200
DD_T=2048 DD_B=256 DD_NDM=2 timeout 600 ncu --set full --clock-control none -k regex:dedisperse_few -s 1 -c 1 -f -o gpurun_out/r2_prof_dd2 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-200
