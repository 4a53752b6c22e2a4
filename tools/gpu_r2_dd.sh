#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -3
: > gpurun_out/r2_dedisp9.jsonl
for blk in 256 1024; do
for cfg in "2048 256 2" "2048 256 3" "2048 256 4" "2048 256 6" "2048 256 8" "4096 32 4" "4096 32 8"; do
  set -- $cfg
  echo -n "blk=$blk "; DSABF_DEDISP_BLOCK=$blk DSABF_DEDISP_FORCE_FEW=1 DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r2_dedisp9.jsonl | cut -c1-100,190-360
done; done
