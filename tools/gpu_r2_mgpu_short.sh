#!/bin/bash
# multi-GPU checkpoint (gpurun --gpus N), tight limits: parity on every rank (tools/mgpu_check.py), then our bench arm at N
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/mgpu_check.py 2>&1 | grep -E "^rank|Error|error|Traceback" | sort | tee gpurun_out/r2_mgpu${N}_check.txt | cut -c1-200 | tail -30
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.4e  sustained %.4e  e2e %.4e (%.2f of copy ceiling)  ms/step %.3f" % (d["n_gpus"], d["value"], d["value_sustained"], d["e2e"]["value"], d["e2e"]["frac_of_copy_ceiling"], d["ms_per_step"]))
x = d["dedisp_full_band"]; print("exchange:", json.dumps(x["exchange"])[:300]); print("dedisperse: %.2f ms %.3e adds/s/gpu | %s" % (x["dedisperse"]["ms"], x["dedisperse"]["adds_per_s_per_gpu"], x["check"]))
PY
