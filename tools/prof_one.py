#!/usr/bin/env python3
"""Launch the fused kernel a few times on one shape (for ncu): prof_one.py obs|dbg|big [exact] [n_launch]
(big = 256 antennas x 2048 beams x 256 channels, 8 chunks: the streamed-weights kernel)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsabeamformer_b200 as bf

shape = sys.argv[1] if len(sys.argv) > 1 else "obs"
exact = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n_ant, n_beams = (256, 2048) if shape == "big" else (64, 256)
B = bf.Beamformer(debug=(shape == "dbg"), exact_power=exact, n_antennas=n_ant, n_beams=n_beams)
rng = np.random.default_rng(1)
B.set_weights(rng.integers(-127, 128, size=(256, n_ant, n_beams, 2), dtype=np.int8))
n_chunks = {"obs": 32, "dbg": 512, "big": 8}[shape]
ts = torch.cuda.Stream()
with torch.cuda.stream(ts):
    x = torch.randint(0, 256, (n_chunks * B.chunk_bytes,), dtype=torch.uint8, device="cuda")
    out = torch.empty(n_chunks * B.chunk_out_floats, dtype=torch.float32, device="cuda")
    for _ in range(n_launch):
        B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
print("done", shape, exact)
