#!/usr/bin/env python
"""Time (and parity-check against the product build's output) every tools/variants/libdsabf_*.so on the headline
shape: one process per variant (DSABF_LIB), 128 MiB block, device-resident, CUDA events, best of 5 x 20 launches."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, hashlib
import numpy as np, torch
sys.path.insert(0, %r)
import dsabeamformer_b200 as bf
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
B = bf.Beamformer(debug=False)
rng = np.random.default_rng(1)
B.set_weights(rng.integers(-127, 128, size=(256, 64, 256, 2), dtype=np.int8))
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randint(0, 256, (32 * B.chunk_bytes,), dtype=torch.uint8, device="cuda", generator=g)
out = torch.zeros(32 * B.chunk_out_floats, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5): B.beamform_device(x.data_ptr(), 32, out.data_ptr(), st)
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): B.beamform_device(x.data_ptr(), 32, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
# DEBUG shape too (G = 2)
B2 = bf.Beamformer(debug=True); B2.set_weights(rng.integers(-127, 128, size=(256, 64, 256, 2), dtype=np.int8))
x2 = torch.randint(0, 256, (2048 * B2.chunk_bytes,), dtype=torch.uint8, device="cuda", generator=g)
o2 = torch.zeros(2048 * B2.chunk_out_floats, dtype=torch.float32, device="cuda")
for _ in range(3): B2.beamform_device(x2.data_ptr(), 2048, o2.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): B2.beamform_device(x2.data_ptr(), 2048, o2.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
h2 = hashlib.sha256(o2.cpu().numpy().tobytes()).hexdigest()[:16]
print(json.dumps(dict(ms=best, tops=2.749e11 / best * 1e-9, sha=h, dbg_ms=e0.elapsed_time(e1) / 10, dbg_sha=h2)))
''' % ROOT
libs = [os.path.join(ROOT, "dsabeamformer_b200", "csrc", "libdsabf.so")] + sorted(l for l in glob.glob(os.path.join(ROOT, "tools", "variants", "libdsabf_*.so")) if "libdsabf_tr_" not in l and "libdsabf_ft_" not in l)
for lib in libs:
    env = dict(os.environ, DSABF_LIB=lib)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])
    except subprocess.TimeoutExpired:
        line = "TIMEOUT"
    print(os.path.basename(lib), line, flush=True)
