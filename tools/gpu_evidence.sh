#!/bin/bash
# final evidence for profiles/: launch list of the bench command, ncu --set full of the two fused kernels and the dedispersion kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/bench_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_kernel -s 2 -c 1 -f -o gpurun_out/prof_obs6 python tools/prof_one.py obs 0 4 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_kernel -s 2 -c 1 -f -o gpurun_out/prof_dbg6 python tools/prof_one.py dbg 0 4 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_big -s 2 -c 1 -f -o gpurun_out/prof_big2 python tools/prof_one.py big 0 4 2>&1 | tail -1
DD_T=2048 DD_B=256 DD_NDM=1 timeout 600 ncu --set full --clock-control none -k regex:dedisperse -s 2 -c 1 -f -o gpurun_out/prof_dd1 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-200
DD_T=512 DD_B=256 DD_NDM=64 timeout 600 ncu --set full --clock-control none -k regex:dedisperse -s 2 -c 1 -f -o gpurun_out/prof_dd64 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-200
