"""include/dsabf_compat.cuh: the names of the reference's CUDA header (src/beamformer.cuh) for source-level
compatibility -- gpuErrchk, gpuBLASchk, START_TIMER / STOP_RECORD_TIMER, expand_input, detect_sum with
detect_dimGrid / detect_dimBlock, print_data*, CUDA_select_GPU -- compile, and on the GPU give what the oracle gives."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "src", "compat_check.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _build(out):
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), SRC, "-o", str(out)],
                          stderr=subprocess.DEVNULL)
    return out


def test_compat_header_compiles(tmp_path):
    """no GPU needed: a translation unit using every compat name builds for sm_100a"""
    exe = _build(tmp_path / "compat_check")
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
def test_compat_kernels_match_oracle(oracle, tmp_path):
    exe = _build(tmp_path / "compat_check")
    rng = np.random.default_rng(11)
    n_bytes = 4 * 256 * 64 * 16                                        # 4 channels x 256 columns x 64 antennas... any multiple of 4
    packed = rng.integers(0, 256, size=n_bytes, dtype=np.uint8)
    packed[:256] = np.arange(256, dtype=np.uint8)                      # every byte value, incl. the reference's known answers
    c = rng.integers(-255, 256, size=(4, 8 * 32, 256, 2)).astype(np.float32)     # [f][o*n_avg + i][beam][re, im], exact in fp32
    packed.tofile(tmp_path / "in.u8"); c.tofile(tmp_path / "c.f32")
    r = subprocess.run([str(exe), str(tmp_path / "in.u8"), str(n_bytes), str(tmp_path / "c.f32"), str(tmp_path / "out")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "compat_check ok" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out.expanded.i8", dtype=np.int8).reshape(-1, 2)
    ref = oracle.unpack(packed.reshape(1, -1, 64)).reshape(-1, 2)
    assert np.array_equal(got, ref)
    for byte, (re, im) in ((0x25, (2, 5)), (0xA8, (-6, -8)), (0xD7, (-3, 7)), (0x70, (7, 0))):      # sandbox/bitshift.cpp, kernelTest.cu
        assert tuple(got[byte]) == (re, im)
    power = np.fromfile(tmp_path / "out.power.f32", dtype=np.float32).reshape(8, 4, 256)
    want = (c.astype(np.float64) ** 2).sum(axis=3).reshape(4, 8, 32, 256).sum(axis=2).transpose(1, 0, 2)
    assert np.array_equal(power.astype(np.float64), want)             # integers below 2^24: exact in any order
