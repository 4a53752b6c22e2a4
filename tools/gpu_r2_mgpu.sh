#!/bin/bash
# multi-GPU checkpoint (gpurun --gpus N): parity on every rank (tools/mgpu_check.py), then both bench arms at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/mgpu_check.py 2>&1 | grep -E "^rank|Error|error|Traceback" | sort | tee gpurun_out/r2_mgpu${N}_check.txt | cut -c1-220 | tail -60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.4e  sustained %.4e  e2e %.4e (%.2f of copy ceiling, h2d %.1f GB/s/GPU, ceiling %.1f)  ms/step %.3f" % (d["n_gpus"], d["value"], d["value_sustained"], d["e2e"]["value"], d["e2e"]["frac_of_copy_ceiling"], d["e2e"]["h2d_gb_s_per_gpu"], d["e2e"]["copy_ceiling"]["h2d_gb_s_per_gpu_with_d2h"], d["ms_per_step"]))
s = d["scaled_sharded"]; print("scaled_sharded: %.3f ms  %.1f TOPS total  frac/gpu %.2f" % (s["ms_per_launch"], s["int8_tops_total"], s["roofline"]["frac"] or 0))
x = d["dedisp_full_band"]; print("exchange:", json.dumps(x["exchange"])[:400]); print("dedisperse: %.2f ms %.3e adds/s/gpu | %s" % (x["dedisperse"]["ms"], x["dedisperse"]["adds_per_s_per_gpu"], x["check"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-200
