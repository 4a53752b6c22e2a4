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
t = tr.reshape(128, 16)
t0 = t[8, 0]
print("stage  mma:wait_start acquired issued | epi warp 8 (slice 0, first of its pair): loop_top loaded math_start math_issued | epi warp 23 (slice 3, second of its pair): ... | unpack(item): start stg_ok b_empty_ok done")
for n in range(40, 72):
    r = t[n] - t0
    it = n // 4
    u = tr.reshape(32, 64)[it % 32][12:16] - t0 if n % 4 == 0 and it < 32 else None
    print(n, r[0:3].tolist(), r[4:8].tolist(), r[8:12].tolist(), u.tolist() if u is not None else "")
d = np.diff(t[40:120, 2])
print("mean stage period (MMA issue to issue):", d.mean(), "clk")
