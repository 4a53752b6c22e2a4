#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "dedisperse" 2>&1 | tail -4
timeout 600 python tools/dedisp_bench.py 2>&1 | tail -1 | tee gpurun_out/dedisp2.json
DD_NDM=64 DD_B=256 DD_T=512 timeout 600 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/dedisp2.json
DD_NDM=8 DD_B=256 DD_T=512 timeout 600 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/dedisp2.json
DD_NDM=1 DD_B=256 DD_T=512 timeout 600 python tools/dedisp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/dedisp2.json
