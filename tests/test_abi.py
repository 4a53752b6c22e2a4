"""CPU: the C-ABI library loads and exports every symbol include/dsabf.h declares;
host-side entry points that need no GPU behave; no oracle in the product."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "dsabf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dsabf_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(bf):
    L = ctypes.CDLL(bf.LIB_PATH)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), n


def test_default_config_matches_reference_constants(bf):
    """src/beamformer.hh:47-60,108-147"""
    obs, dbg = bf.default_config(False), bf.default_config(True)
    L = bf.lib()
    for c, navg, cols, chunk, block in ((obs, 16, 256, 4194304, 134217728), (dbg, 1, 16, 262144, 8388608)):
        assert (c.n_antennas, c.n_beams, c.n_channels, c.n_pol, c.n_averaging) == (64, 256, 256, 2, navg)
        assert (c.n_outputs_per_chunk, c.n_chunks_per_block, c.n_ring_slots) == (8, 32, 8)
        assert L.dsabf_columns_per_chunk(ctypes.byref(c)) == cols
        assert L.dsabf_chunk_bytes(ctypes.byref(c)) == chunk
        assert L.dsabf_block_bytes(ctypes.byref(c)) == block
        assert L.dsabf_chunk_out_floats(ctypes.byref(c)) == 8 * 256 * 256


def test_host_weights_match_reference_golden(bf, oracle, golden, config_dir):
    """dsabf_compute_weights (product host code) == the reference's compiled loop, bit for bit."""
    import hashlib
    g, _ = golden
    for case in ("linear_g0", "grid3721_g3", "defaults_bogus_g5"):
        e = g[case]
        pos = oracle.read_positions(os.path.join(config_dir, e["positions"])) if e["positions"] else oracle.default_positions()
        dirs = oracle.read_directions(os.path.join(config_dir, e["directions"]), 256) if e["directions"] else oracle.default_directions()
        w = bf.compute_weights(pos, dirs, e["gpu"])
        assert hashlib.sha256(w.tobytes()).hexdigest() == e["weights_sha256"], case


def test_create_without_gpu_fails_loudly(bf):
    import torch
    if torch.cuda.is_available():
        return
    try:
        bf.Beamformer()
    except bf.DsabfError as e:
        assert e.code != 0
    else:
        raise AssertionError("dsabf_create must fail without a GPU (no CPU fallback)")


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dsabeamformer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, f)


def test_dm_plan_matches_oracle(oracle):
    """dsabf_dm_trials / dsabf_dm_delays (product host code, no GPU needed) == the oracle's restatement of
    sandbox/Dispersion Theory.ipynb cells 2 and 5: 1627 trials to DM 2000, identical integer delays over the band."""
    import dsabeamformer_b200 as bf
    dms = bf.dm_trials(2000.0, 2048)
    ref = oracle.dm_trials(2000.0, 2048)
    assert len(dms) == len(ref) == 1627 and np.array_equal(dms, ref)
    assert len(bf.dm_trials(100.0, 840)) == len(oracle.dm_trials(100.0, 840))
    freq = 1.28 + (1.53 - 1.28) / 2048 * np.arange(2048)               # the notebook's channel grid
    d = bf.dm_delays(dms[::25], freq)
    assert d.dtype == np.int32 and np.array_equal(d, oracle.dm_delays(dms[::25], freq))
    assert d[0].max() == 0 and d[-1, 0] == int(4.15 * dms[::25][-1] * (1.28 ** -2 - 1.53 ** -2) / (0.131 * 16))


def test_experiment_patches_still_apply():
    """tools/experiments/ keeps the dropped kernel variants as patches against the product source; they must not rot
    silently (checked only in a git checkout: the GPU box gets a snapshot without .git)."""
    import shutil
    import subprocess
    if not os.path.isdir(os.path.join(ROOT, ".git")) or not shutil.which("git"):
        pytest.skip("not a git checkout")
    exp = os.path.join(ROOT, "tools", "experiments")
    for patch, reverse in (("r02_mix_and_turns.patch", False), ("r02_epilogue_switches_removed.patch", True)):
        cmd = ["git", "apply", "--check"] + (["-R"] if reverse else []) + [os.path.join(exp, patch)]
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, (patch, r.stderr[-400:])
