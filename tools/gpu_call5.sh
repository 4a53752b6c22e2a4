#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_r01_a.json | tail -3
