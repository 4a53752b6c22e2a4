#!/bin/bash
for shape in "64 256" "256 64"; do
  echo "== $shape"; DSABF_FORCE_STREAMED=1 timeout 40 python tools/dbg_big.py $shape 2>&1 | tail -2 | cut -c1-200
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or scaled" 2>&1 | tail -15
DSABF_FORCE_STREAMED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
