"""Clock64 trace of ONE stage-owning epilogue warp and of both MMA issuers (library built with -DDSABF_TRACE=2
[-DDSABF_TRACE_EW=n]; DSABF_LIB selects it).  Stamps per visit: top, full_ok, then per 64-column half: loaded,
(arrived), block0 done, block1 done."""
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
rows = [a for a in range(24, 120) if T[a, 4] != 0]
t = T[rows, 4:14]
names = ["top->full_ok", "->h0 loaded", "->(h0 index)", "->h0 block0", "->h0 block1", "->h1 loaded", "->h1 arrived", "->h1 block0", "->h1 block1"]
d = np.diff(t, axis=1)
print("lib =", os.environ.get("DSABF_LIB", "default"), " rows traced:", rows[:4], "...")
for i in range(9):
    print("  %-18s %7.1f  (min %d max %d)" % (names[i], d[:, i].mean(), d[:, i].min(), d[:, i].max()))
nxt = t[1:, 0] - t[:-1, 9]
print("  %-18s %7.1f  (min %d max %d)" % ("end->next top", nxt.mean(), nxt.min(), nxt.max()))
print("  visit period       %7.1f" % np.diff(t[:, 0]).mean())
m = T[40:56, 0:3]; base = m[0, 0]
print("MMA issuers, stages 40..55: wait_start acquired issued")
for a in range(16): print("  acc %d: %s   full_ok(epi) %s  arrived %s" % (40 + a, (m[a] - base).tolist(), (T[40 + a, 5] - base) if T[40 + a, 5] else "-", (T[40 + a, 11] - base) if T[40 + a, 11] else "-"))
