import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dsabeamformer_b200 as bf
n_ant, n_beams = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
B = bf.Beamformer(debug=False, n_channels=2, n_antennas=n_ant, n_beams=n_beams)
B.set_weights(rng.integers(-127, 128, size=(2, n_ant, n_beams, 2), dtype=np.int8))
x = rng.integers(0, 256, size=(1, 2, B.cols_per_chunk, n_ant), dtype=np.uint8)
out = B.beamform(x)
print("ok", out.shape, float(out.sum()))
