#!/bin/bash
# dedispersion: parity tests, then the few-trials streaming kernel at 5..8 trials over block size x channels per step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -3
: > gpurun_out/r2_dedisp11.jsonl
for un in 2 4; do
for blk in 256 512 1024; do
for cfg in "2048 256 4" "2048 256 5" "2048 256 6" "2048 256 7" "2048 256 8"; do
  set -- $cfg
  echo -n "unroll=$un blk=$blk "; DSABF_DEDISP_UNROLL=$un DSABF_DEDISP_BLOCK=$blk DSABF_DEDISP_FORCE_FEW=1 DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r2_dedisp11.jsonl | cut -c1-100,190-380
done; done; done
