#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants6.txt 2>&1; cat gpurun_out/r2_variants6.txt | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -4
: > gpurun_out/r2_dedisp6.jsonl
for cfg in "512 256 2" "512 256 4" "512 256 8" "2048 256 2" "2048 256 4" "2048 256 8" "4096 32 2" "4096 32 8"; do
  set -- $cfg
  DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r2_dedisp6.jsonl | cut -c1-100,190-360
done
