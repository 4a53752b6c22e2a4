#!/bin/bash
mkdir -p gpurun_out
DD_T=2048 DD_B=256 DD_NDM=1 timeout 600 ncu --set full --clock-control none -k regex:dedisperse -s 2 -c 1 -f -o gpurun_out/prof_dd1b python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-120
DD_T=512 DD_B=256 DD_NDM=64 timeout 600 ncu --set full --clock-control none -k regex:dedisperse -s 2 -c 1 -f -o gpurun_out/prof_dd64b python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-120
