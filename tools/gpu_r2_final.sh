#!/bin/bash
# end-of-round evidence on one B200: full -m gpu suite, smoke(), bench, ncu launch list of the bench command (tight time limits:
# a hung step must not eat the GPU budget)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_gputests_final.txt
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke_final.txt
timeout 420 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err; cut -c1-400 gpurun_out/r2_bench_final.json
if [ "$1" = "full" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_final.json; cut -c1-200 gpurun_out/r2_bench_reference_final.json
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --sustain 0 > gpurun_out/r2_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_bench_launches.csv
fi
