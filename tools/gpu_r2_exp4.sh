#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants4.txt 2>&1
for v in r1 r1hi; do DSABF_LIB=$PWD/tools/variants/libdsabf_ft_$v.so timeout 200 python tools/trace_fine.py > gpurun_out/r2_ftrace_$v.txt 2>&1; done
cat gpurun_out/r2_variants4.txt gpurun_out/r2_ftrace_r1.txt gpurun_out/r2_ftrace_r1hi.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_gputests4.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_try1.json 2> gpurun_out/r2_bench_try1.err; tail -3 gpurun_out/r2_bench_try1.err; cut -c1-1500 gpurun_out/r2_bench_try1.json
