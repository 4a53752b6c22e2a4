#!/usr/bin/env python
"""Time the fused kernel on an arbitrary shape (device-resident input, CUDA events).
usage: time_shape.py n_ant n_beams n_chan n_chunks [exact] [reps]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import dsabeamformer_b200 as bf


def time_shape(n_ant, n_beams, n_chan, n_chunks, exact=0, reps=10, debug=False):
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    B = bf.Beamformer(debug=debug, n_antennas=n_ant, n_beams=n_beams, n_channels=n_chan, exact_power=exact)
    rng = np.random.default_rng(1)
    B.set_weights(rng.integers(-127, 128, size=(n_chan, n_ant, n_beams, 2), dtype=np.int8))
    x = torch.randint(0, 256, (n_chunks * B.chunk_bytes,), dtype=torch.uint8, device="cuda")
    out = torch.empty(n_chunks * B.chunk_out_floats, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cols = n_chunks * B.cols_per_chunk
    ops = 8.0 * n_ant * n_beams * n_chan * cols            # 4 real MACs per complex MAC
    samples = n_beams * n_chan * cols / 2.0                # beams x chans x time (2 pol columns per time)
    byts = n_chunks * (B.chunk_bytes + 4 * B.chunk_out_floats)
    r = dict(n_ant=n_ant, n_beams=n_beams, n_chan=n_chan, n_chunks=n_chunks, cols=cols, exact=exact, ms=ms,
             tops=ops / ms * 1e-9, samples_per_s=samples / ms * 1e3, hbm_gbs=byts / ms * 1e-6)
    B.close()
    return r


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    r = time_shape(a[0], a[1], a[2], a[3], a[4] if len(a) > 4 else 0, a[5] if len(a) > 5 else 10)
    print(json.dumps(r))
