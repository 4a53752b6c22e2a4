#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -3
: > gpurun_out/r2_dedisp7.jsonl
for cfg in "512 256 2" "512 256 4" "2048 256 2" "2048 256 3" "2048 256 4" "4096 32 2" "4096 32 4"; do
  set -- $cfg
  DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r2_dedisp7.jsonl | cut -c1-100,190-360
done
DSABF_DEDISP_FORCE_FEW=1 DD_T=2048 DD_B=256 DD_NDM=4 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-100,190-360
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_beamform_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_obs python tools/prof_one.py obs 0 4 2>&1 | tail -1
