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
