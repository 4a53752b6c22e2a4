#!/bin/bash
# first GPU call: microbenchmarks, first-light parity, reference kernels sanity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== ubench" ; timeout 300 ./tools/ubench 2>&1 | tee gpurun_out/ubench.txt
echo "=== first light" ; timeout 900 python tools/first_light.py 2>&1 | tee gpurun_out/first_light.txt
echo "=== reference kernels (sm_100 rebuild)"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/ref_kernels.txt
import sys, os, subprocess, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tools')
import oracle as O, make_configs
make_configs.make_all('config')
pos = O.read_positions('config/linear_positions.txt'); dirs = O.read_directions('config/linear_directions.txt', 256)
src = O.read_directions('config/linear_source_directions_1024.txt')
w = O.weights(pos, dirs, 0); w.tofile('/tmp/w.i8')
for tag, cols, navg in (("dbg", 16, 1), ("obs", 256, 16)):
    d = O.generate(pos, src, batch=0, n_chunks=1024, cols=cols)[[100, 511, 900]] if cols == 16 else O.generate(pos, src[[100, 511, 900]], batch=0, n_chunks=3, cols=cols)
    d.tofile('/tmp/chunks_%s.u8' % tag)
    r = subprocess.run(['oracle/_ref/ref_kernels_%s' % tag, '/tmp/w.i8', '/tmp/chunks_%s.u8' % tag, '3', '/tmp/ref_%s' % tag, '20'], capture_output=True, text=True)
    print(tag, r.stdout, r.stderr[-500:])
    C = np.fromfile('/tmp/ref_%s.C.f32' % tag, dtype=np.float32).reshape(256, cols, 256, 2)
    out = np.fromfile('/tmp/ref_%s.out.f32' % tag, dtype=np.float32).reshape(3, 8, 256, 256)
    V = O.beamform(d[0], w)
    alpha = np.float32(1.0 / 127)
    Cexp = (V.astype(np.float32) * alpha)
    print(tag, "cuBLAS C == fl(alpha*V):", np.array_equal(C, Cexp), "max|C|", np.abs(C).max(), "nonzero", np.count_nonzero(C))
    G = 2 * navg
    for c in range(3):
        Vc = O.beamform(d[c], w)
        Fo = O.detect_ref_float(Vc, G)
        print(tag, "chunk", c, "detect_sum == oracle float emulation:", np.array_equal(out[c], Fo), "max rel", (np.abs(out[c] - Fo) / np.maximum(Fo, 1e-30)).max())
PY
