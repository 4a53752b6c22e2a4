#!/usr/bin/env python3
"""GPU bring-up script: parity of the fused kernel against the oracle on small and
full shapes, then a first timing.  Run on the GPU box: python tools/first_light.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import dsabeamformer_b200 as bf  # noqa: E402


def check(name, debug, n_chan, n_chunks, n_beams=256, seed=0, exact=1, **kw):
    rng = np.random.default_rng(seed)
    B = bf.Beamformer(debug=debug, n_channels=n_chan, n_beams=n_beams, exact_power=exact, **kw)
    w = rng.integers(-127, 128, size=(n_chan, 64, n_beams, 2), dtype=np.int8)
    B.set_weights(w)
    cols = B.cols_per_chunk
    packed = rng.integers(0, 256, size=(n_chunks, n_chan, cols, 64), dtype=np.uint8)
    t = time.time()
    V, P = B.voltages(packed)
    out = B.beamform(packed)
    tg = time.time() - t
    # oracle: per chunk
    ok = True
    G = B.group
    n_out = cols // G
    for c in range(n_chunks):
        Vo = O.beamform(packed[c], w)                      # [chan][cols][beam][2]
        Po = O.power_int(Vo, G)                            # [n_out][chan][beam]
        Fo = O.detect_ref_float(Vo, G)
        Vg = V[:, c * cols:(c + 1) * cols]
        Pg = P[c * n_out:(c + 1) * n_out]
        Fg = out[c * n_out:(c + 1) * n_out]
        v_ok = np.array_equal(Vo, Vg)
        p_ok = np.array_equal(Po, Pg)
        rel = np.abs(Fg.astype(np.float64) - Fo) / np.maximum(np.abs(Fo), 1e-30)
        a2 = float(np.float32(1.0 / 127)) ** 2
        exact_rel = np.abs(Fg.astype(np.float64) - Po * a2) / np.maximum(Po * a2, 1e-30)
        f_ok = rel.max() < 1e-6
        if not (v_ok and p_ok and f_ok) or c == 0:
            print("  %s chunk %d: V %s (ndiff %d) P %s float-vs-ref max rel %.2e (vs exact %.2e)" % (
                name, c, v_ok, int((Vo != Vg).sum()), p_ok, rel.max(), exact_rel.max()))
        ok &= v_ok and p_ok and f_ok
    print("%s: %s  (n_chan %d n_chunks %d cols %d G %d beams %d exact %d) gpu-path %.2fs launches %d" % (
        name, "PASS" if ok else "FAIL", n_chan, n_chunks, cols, G, n_beams, exact, tg, B.kernel_launches))
    B.close()
    return ok


def timing():
    import torch
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    B = bf.Beamformer(debug=False)
    rng = np.random.default_rng(1)
    w = rng.integers(-127, 128, size=(256, 64, 256, 2), dtype=np.int8)
    B.set_weights(w)
    for exact in (1, 0):
        B2 = bf.Beamformer(debug=False, exact_power=exact)
        B2.set_weights(w)
        n_chunks = 32
        x = torch.randint(0, 256, (n_chunks * B2.chunk_bytes,), dtype=torch.uint8, device="cuda")
        out = torch.empty(n_chunks * B2.chunk_out_floats, dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            B2.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            B2.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        samples = 256 * 256 * n_chunks * 8 * 16
        print("obs-shape block (32 chunks, 128 MiB) exact=%d: %.3f ms/launch  %.3e samples/s  %.1f TOPS int8  in %.1f GB/s" % (
            exact, ms, samples / ms * 1e3, samples * 1024 / ms * 1e-9, n_chunks * B2.chunk_bytes / ms * 1e-6))
        B2.close()
    Bd = bf.Beamformer(debug=True)
    Bd.set_weights(w)
    n_chunks = 1024
    x = torch.randint(0, 256, (n_chunks * Bd.chunk_bytes,), dtype=torch.uint8, device="cuda")
    out = torch.empty(n_chunks * Bd.chunk_out_floats, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        Bd.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        Bd.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    samples = 256 * 256 * n_chunks * 8
    byts = n_chunks * (Bd.chunk_bytes + Bd.chunk_out_floats * 4)
    print("debug-shape 1024 chunks: %.3f ms/launch  %.3e samples/s  %.1f TOPS  HBM %.1f GB/s" % (
        ms, samples / ms * 1e3, samples * 1024 / ms * 1e-9, byts / ms * 1e-6))


if __name__ == "__main__":
    if "--time-only" in sys.argv:
        timing()
        sys.exit(0)
    allok = True
    allok &= check("tiny-obs", False, 2, 1, seed=1)
    allok &= check("tiny-dbg", True, 3, 32, seed=2)
    allok &= check("dbg-partial-tile", True, 2, 21, seed=3)
    allok &= check("obs-64beams", False, 2, 2, n_beams=64, seed=4)
    allok &= check("obs-float-epilogue", False, 2, 2, seed=5, exact=0)
    allok &= check("navg4", False, 5, 3, seed=6, n_averaging=4)
    allok &= check("many-chan-obs", False, 300, 1, seed=7)
    print("ALL", "PASS" if allok else "FAIL")
    if allok or "--time" in sys.argv:
        timing()
