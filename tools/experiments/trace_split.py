"""Clock64 trace of the split-N kernel: one epilogue warp (DSABF_TRACE_EW) and both MMA issuers."""
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
T = tr.reshape(64, 2, 16)
g = 0 if T[30, 0, 4] != 0 else 1
t = T[12:60, g, 4:9]
names = ["top->full_ok", "->loaded", "->arrived", "->math+store done"]
d = np.diff(t, axis=1)
print("lib =", os.environ.get("DSABF_LIB", "default"), " traced epilogue warp is in group", g)
for i in range(4):
    print("  %-22s %7.1f  (min %d max %d)" % (names[i], d[:, i].mean(), d[:, i].min(), d[:, i].max()))
nxt = t[1:, 0] - t[:-1, 4]
print("  %-22s %7.1f  (min %d max %d)" % ("end->next top", nxt.mean(), nxt.min(), nxt.max()))
print("  period                 %7.1f" % np.diff(t[:, 0]).mean())
base = T[20, 0, 0]
print("MMA issuers, accumulators 20..31: group 0 [wait_start acquired issued] | group 1 [...] | traced warp: top, loaded, end")
for a in range(20, 32): print("  acc %d: %s | %s | %s" % (a, (T[a, 0, 0:3] - base).tolist(), (T[a, 1, 0:3] - base).tolist(), (T[a, g, [4, 6, 8]] - base).tolist()))
