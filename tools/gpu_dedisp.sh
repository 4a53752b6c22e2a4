#!/bin/bash
# dedispersion: parity, then timing over the trial-count regimes (HBM-bound -> shared-memory-bound)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dedisperse or dm0" 2>&1 | tail -3
: > gpurun_out/dedisp.jsonl
for cfg in "512 256 1" "2048 256 1" "1024 32 1" "512 256 2" "512 256 4" "512 256 8" "512 256 64" "1024 32 256" "1024 256 256"; do
  set -- $cfg
  DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/dedisp.jsonl | cut -c1-100,190-340
done
