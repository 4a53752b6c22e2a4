#!/bin/bash
mkdir -p gpurun_out
echo "=== first light" ; timeout 900 python tools/first_light.py 2>&1 | tee gpurun_out/first_light.txt
