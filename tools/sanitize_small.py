"""Tiny fused-kernel + dedispersion run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dsabeamformer_b200 as bf
rng = np.random.default_rng(0)
for debug, n_chan, n_chunks in ((False, 2, 1), (True, 2, 21)):
    B = bf.Beamformer(debug=debug, n_channels=n_chan)
    B.set_weights(rng.integers(-127, 128, size=(n_chan, 64, 256, 2), dtype=np.int8))
    x = rng.integers(0, 256, size=(n_chunks, n_chan, B.cols_per_chunk, 64), dtype=np.uint8)
    out = B.beamform(x)
    V, P = B.voltages(x)
    print("ok", debug, float(out.sum()), int(P.sum() % 1000003))
    if debug:
        d = B.dedisperse(np.ascontiguousarray(out[:40, :, :64]), rng.integers(0, 9, size=(3, n_chan)).astype(np.int32), 30)
        print("dedisp", float(d.sum()))
    B.close()
# streamed-weights kernel (two work items per CTA at most, 4 k-blocks, odd number of M-tiles) and the few-trial
# dedispersion variants (warps split the output times; 4-deep pipeline)
B = bf.Beamformer(debug=False, n_channels=3, n_antennas=256, n_beams=192)
B.set_weights(rng.integers(-127, 128, size=(3, 256, 192, 2), dtype=np.int8))
x = rng.integers(0, 256, size=(1, 3, B.cols_per_chunk, 256), dtype=np.uint8)
out = B.beamform(x)
V, P = B.voltages(x)
print("ok streamed", float(out.sum()), int(P.sum() % 1000003))
pw = rng.random((80, 9, 40), dtype=np.float32)
for n_dm in (1, 2, 4, 8):
    d = B.dedisperse(pw, rng.integers(0, 9, size=(n_dm, 9)).astype(np.int32), 60)
    print("dedisp", n_dm, float(d.sum()))
B.close()
