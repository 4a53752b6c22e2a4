#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_ddvar.txt
for lib in dsabeamformer_b200/csrc/libdsabf.so tools/variants/libdsabf_dd*.so; do
  echo "== $lib" | tee -a gpurun_out/r2_ddvar.txt
  DSABF_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -1 | tee -a gpurun_out/r2_ddvar.txt
  for cfg in "512 256 64" "1024 256 256" "2048 32 1627"; do
    set -- $cfg
    DSABF_LIB=$PWD/$lib DD_T=$1 DD_B=$2 DD_NDM=$3 timeout 300 python tools/dedisp_bench.py 2>&1 | tail -1 | cut -c1-100,190-330 | tee -a gpurun_out/r2_ddvar.txt
  done
done
