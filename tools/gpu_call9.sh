#!/bin/bash
# round-1 checkpoint: full GPU suite, bench line, launch list, ncu captures of both fused kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 1500 gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_kernel -s 2 -c 1 -f -o gpurun_out/prof_obs5 python tools/prof_one.py obs 0 4 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_big -s 2 -c 1 -f -o gpurun_out/prof_big1 python tools/prof_one.py big 0 4 2>&1 | tail -1
