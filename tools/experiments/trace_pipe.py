"""Clock64 trace of ONE chunk-pipelined epilogue warp and of both MMA issuers (library built with -DDSABF_TRACE=2
[-DDSABF_TRACE_EW=n]; DSABF_LIB selects it)."""
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
T = tr.reshape(128, 16)
t = T[24:120, 4:11]
names = ["top->A ready, ld B issued", "->math A done", "->B ready", "->arrived, probed, (ld A')", "->math B done + store", "->fallback wait + ld A'"]
d = np.diff(t, axis=1)
print("lib =", os.environ.get("DSABF_LIB", "default"))
for i in range(6):
    print("  %-30s %7.1f  (min %d max %d)" % (names[i], d[:, i].mean(), d[:, i].min(), d[:, i].max()))
nxt = t[1:, 0] - t[:-1, 6]
print("  %-30s %7.1f  (min %d max %d)" % ("end->next top", nxt.mean(), nxt.min(), nxt.max()))
print("  period                         %7.1f   prefetched next chunk 0 early in %.0f %% of the stages" % (np.diff(t[:, 0]).mean(), 100 * T[24:120, 11].mean()))
m = T[40:52, 0:3]; base = m[0, 0]
print("MMA issuers, stages 40..51: wait_start acquired issued | epilogue top, B ready(all loaded), end")
for a in range(12): print("  acc %d: %s | %s" % (40 + a, (m[a] - base).tolist(), (T[40 + a, [4, 7, 10]] - base).tolist()))
