#!/bin/bash
mkdir -p gpurun_out
for v in base base15 pp0 pp0b; do
  DSABF_LIB=$PWD/tools/variants/libdsabf_ft_$v.so timeout 200 python tools/trace_fine.py > gpurun_out/r2_ftrace_$v.txt 2>&1
done
cat gpurun_out/r2_ftrace_*.txt
