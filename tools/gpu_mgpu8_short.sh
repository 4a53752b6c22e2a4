#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 tools/mgpu_check.py 2>&1 | grep -E "^rank|Error|error" | sort > gpurun_out/mgpu8_check.txt; grep -c "True\|max rel" gpurun_out/mgpu8_check.txt; grep "rank 0 timing" gpurun_out/mgpu8_check.txt | cut -c1-250
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_n8.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r01_n8.json"))
print("N=%d value %.4e samples/s  e2e %.4e  ms/step %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
