#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -s 2>&1 | tail -12
for n in 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_n$n.json | cut -c1-400
done
