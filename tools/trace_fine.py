"""Fine clock64 trace of ONE epilogue warp (library built with -DDSABF_TRACE=2 [-DDSABF_TRACE_EW=n]; DSABF_LIB selects it):
eight stamps per accumulator stage.  Prints the mean interval between consecutive stamps over stages 24..119."""
import sys, os, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, dsabeamformer_b200 as bf
B = bf.Beamformer(debug=False)
B.set_weights(np.random.default_rng(1).integers(-127, 128, size=(256, 64, 256, 2), dtype=np.int8))
x = torch.randint(0, 256, (32 * B.chunk_bytes,), dtype=torch.uint8, device="cuda")
out = torch.empty(32 * B.chunk_out_floats, dtype=torch.float32, device="cuda")
L = bf.lib(); L.dsabf_debug_trace.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 2
tr = np.zeros(2048, dtype=np.int64)
for _ in range(2):
    rc = L.dsabf_debug_trace(B._h, ctypes.c_void_p(x.data_ptr()), 32, ctypes.c_void_p(out.data_ptr()), tr.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
t = tr.reshape(128, 16)[:, 4:12]
names = ["top->full_ok", "full_ok->loaded", "loaded->fence+syncwarp", "->arrive(acc_empty)", "->math_start", "block0 math", "block1 math (+ shuffle, store)",
         "end->next top"]
sel = t[24:120]
d = np.diff(sel, axis=1)
nxt = t[25:121, 0] - t[24:120, 7]
print("mean clocks per phase (stages 24..119), lib =", os.environ.get("DSABF_LIB", "default"))
for i in range(7):
    print("  %-30s %7.1f  (min %d max %d)" % (names[i], d[:, i].mean(), d[:, i].min(), d[:, i].max()))
print("  %-30s %7.1f  (min %d max %d)" % (names[7], nxt.mean(), nxt.min(), nxt.max()))
print("  period (top to top)            %7.1f" % np.diff(t[24:121, 0]).mean())
print("MMA issuer: wait_start acquired issued (first rows):")
m = tr.reshape(128, 16)[40:56, 0:3]
T = tr.reshape(128, 16)
for a in range(40, 56):
    print("  acc %d: issuer %s | epilogue warp: top %d full_ok %d loaded %d arrived %d end %d" % (a, (T[a, 0:3] - m[0, 0]).tolist(),
          T[a, 4] - m[0, 0], T[a, 5] - m[0, 0], T[a, 6] - m[0, 0], T[a, 8] - m[0, 0], T[a, 11] - m[0, 0]))
