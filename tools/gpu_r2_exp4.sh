#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants4.txt 2>&1
DSABF_LIB=$PWD/tools/variants/libdsabf_ft_r1.so timeout 200 python tools/trace_fine.py > gpurun_out/r2_ftrace_r1.txt 2>&1
cat gpurun_out/r2_variants4.txt gpurun_out/r2_ftrace_r1.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_gputests4.txt
