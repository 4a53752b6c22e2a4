#!/bin/bash
# end-of-round check: full GPU suite, smoke, bench line, dedispersion sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r01.json"))
print("value %.4e  e2e %.4e  kernel_ms %.4f  achieved %.1f frac %.3f  launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["gpu_launches"]))
for w in d["other_workloads"]: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in w.items() if k != "workload"})
print(d["clocks"])
PY
bash tools/gpu_dedisp.sh 2>&1 | tail -10
