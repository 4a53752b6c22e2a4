#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/first_light.py --time-only 2>&1 | tee gpurun_out/timing2.txt
