#!/usr/bin/env python3
"""Run the REFERENCE's kernels + cuBLAS (oracle/_ref/ref_kernels_{dbg,obs}) on a B200 and
save a small fixture of their outputs to gpurun_out/ref_gpu_fixture.npz (copy it to
tests/golden/).  Inputs: linear config, sources 0, 511, 1023, sub-band 0."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_configs  # noqa: E402
import oracle as O  # noqa: E402

cfg = make_configs.make_all(os.path.join(ROOT, "config"))
pos = O.read_positions(os.path.join(cfg, "linear_positions.txt"))
dirs = O.read_directions(os.path.join(cfg, "linear_directions.txt"), 256)
src = O.read_directions(os.path.join(cfg, "linear_source_directions_1024.txt"))
pick = [0, 511, 1023]
w = O.weights(pos, dirs, 0)
fx = {"sources": np.array(pick), "chan_step": 32, "beam_step": 16}
with tempfile.TemporaryDirectory() as tmp:
    w.tofile(os.path.join(tmp, "w.i8"))
    for shape, cols in (("dbg", 16), ("obs", 256)):
        data = O.generate(pos, src[pick], n_chunks=len(pick), cols=cols)
        data.tofile(os.path.join(tmp, "x.u8"))
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_kernels_" + shape), os.path.join(tmp, "w.i8"),
                            os.path.join(tmp, "x.u8"), str(len(pick)), os.path.join(tmp, "ref")], capture_output=True, text=True)
        print(shape, r.stdout.strip(), r.stderr.strip())
        assert r.returncode == 0 and "gemm_status 0" in r.stdout
        out = np.fromfile(os.path.join(tmp, "ref.out.f32"), dtype=np.float32).reshape(len(pick) * 8, 256, 256)
        fx[shape + "_out_sub"] = out[:, ::32, ::16].copy()
        fx[shape + "_ded"] = np.fromfile(os.path.join(tmp, "ref.ded.f32"), dtype=np.float32).reshape(len(pick), 256)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_gpu_fixture.npz"), **fx)
print("wrote gpurun_out/ref_gpu_fixture.npz")
