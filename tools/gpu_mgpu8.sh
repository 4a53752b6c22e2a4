#!/bin/bash
# 8-GPU checkpoint: parity under torchrun on all ranks, then the bench line at N = 8, 4, 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/mgpu_check.py 2>&1 | grep -E "^rank|Error|error" | sort | tee gpurun_out/mgpu8_check.txt | tail -26
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_n$n.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r01_n$n.json"))
print("N=%d value %.4e samples/s  e2e %.4e  ms/step %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
done
