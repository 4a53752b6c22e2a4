#!/bin/bash
# ncu --set full of the two dedispersion kernels changed in the third pass + the bench launch list (tight limits)
mkdir -p gpurun_out
DSABF_DEDISP_FORCE_FEW=1 DD_T=2048 DD_B=256 DD_NDM=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:dedisperse_few -s 2 -c 1 -f -o gpurun_out/r2_prof_ddfew4 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-120
DD_T=2048 DD_B=32 DD_NDM=1627 timeout 240 ncu --set full --clock-control none --import-source on -k regex:dedisperse_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_dd1627b python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-120
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --sustain 0 > gpurun_out/r2_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_bench_launches.csv
ls -la gpurun_out/*.ncu-rep | tail -3
