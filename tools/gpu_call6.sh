#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
timeout 600 python tools/dedisp_bench.py 2>&1 | tail -2 | tee gpurun_out/dedisp1.json
DD_NDM=64 DD_B=256 DD_T=512 timeout 600 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/dedisp1.json
