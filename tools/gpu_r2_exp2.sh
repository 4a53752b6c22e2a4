#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants2.txt 2>&1
for v in base pp0 p1 pp1; do
  DSABF_LIB=$PWD/tools/variants/libdsabf_tr_$v.so timeout 200 python tools/trace_one.py > gpurun_out/r2_trace_$v.txt 2>&1
done
cat gpurun_out/r2_variants2.txt; tail -3 gpurun_out/r2_trace_*.txt
