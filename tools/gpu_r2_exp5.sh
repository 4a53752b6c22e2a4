#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py 2>&1 | tail -15 | tee gpurun_out/r2_gputests5.txt
: > gpurun_out/r2_dedisp.jsonl
for cfg in "512 256 1" "512 256 2" "512 256 4" "512 256 8" "2048 256 2" "2048 256 8" "512 256 64" "1024 32 256" "2048 32 1627"; do
  set -- $cfg
  DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r2_dedisp.jsonl | cut -c1-100,190-360
done
