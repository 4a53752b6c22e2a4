#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_exp1_smi.txt
./tools/ubench_pipes > gpurun_out/r2_ubench_pipes.txt 2>&1
timeout 600 python tools/time_variants.py > gpurun_out/r2_variants.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/r2_parity.txt
cat gpurun_out/r2_ubench_pipes.txt gpurun_out/r2_variants.txt gpurun_out/r2_parity.txt
