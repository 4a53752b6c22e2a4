#!/bin/bash
# time every tools/variants/libdsabf_*.so against the product build on the headline shape (tools/time_variants.py)
mkdir -p gpurun_out
timeout 900 python tools/time_variants.py > gpurun_out/r2_variants_s3.txt 2>&1
cut -c1-200 gpurun_out/r2_variants_s3.txt
