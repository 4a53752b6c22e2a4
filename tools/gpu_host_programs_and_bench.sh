#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_programs.py -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r01.json"))
print("value %.4e  e2e %.4e  kernel_ms %.4f  achieved %.1f frac %.3f  binding %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"].get("host_binding")))
for w in d["other_workloads"]: print(w)
print(d["clocks"])
PY
tail -3 gpurun_out/bench_r01.err
