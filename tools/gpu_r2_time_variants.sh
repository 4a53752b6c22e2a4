#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants8.txt 2>&1; cut -c1-200 gpurun_out/r2_variants8.txt
