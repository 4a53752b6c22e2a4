"""The C++ host programs (bin/beam_debug, bin/beam, bin/ringdb, bin/junkdb): the reference's
command line, config files, stdout milestones and bin/data.py, with the GPU work behind the C ABI."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "bin")


def _need(exe):
    p = os.path.join(BIN, exe)
    if not os.path.exists(p):
        pytest.fail("%s not built: run `make -C dsabeamformer_b200/host` (or __graft_entry__.build())" % p)
    return p


def _load_data_py(path):
    ns = {}
    exec(open(path).read(), ns)
    return np.array(ns["A"], dtype=np.float64)


def test_usage_text_lists_reference_options():
    """-h prints the reference's option list (src/beamformer.hh:222-243) and exits 0 without a GPU."""
    out = subprocess.run([_need("beam_debug"), "-h"], capture_output=True, text=True)
    assert out.returncode == 0
    for opt in ("-g gpu", "-p position_filename", "-d direction_filename", "-s source_filename", "-h "):
        assert opt in out.stdout
    out = subprocess.run([_need("beam"), "-h"], capture_output=True, text=True)
    assert out.returncode == 0
    for opt in ("-c core", "-k key", "-g gpu", "-p position_filename", "-d direction_filename"):
        assert opt in out.stdout


def test_ring_tools_roundtrip(tmp_path):
    """ringdb + junkdb (stand-ins for dada_db / dada_junkdb): blocks arrive in order, full size, then EOD."""
    reader = tmp_path / "reader.cc"
    reader.write_text('''
#include "dada_ring.hh"
#include <cstdio>
int main(int argc, char** argv) {
    unsigned key; sscanf(argv[1], "%x", &key);
    beam::DadaRing r; if (!r.connect(key)) return 2;
    uint64_t hs; r.read_header(&hs);
    uint64_t n = 0, bytes, id; unsigned long long sum = 0;
    for (;;) { char* b = r.open_block_read(&bytes, &id); if (bytes == 0) break; if (id != n) return 3;
        for (uint64_t i = 0; i < bytes; i += 4096) sum += (unsigned char)b[i]; n++; r.close_block_read(); }
    printf("blocks %llu bufsz %llu sum %llu\\n", (unsigned long long)n, (unsigned long long)r.bufsz(), sum); return 0; }
''')
    exe = tmp_path / "reader"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "dsabeamformer_b200", "host"), str(reader), "-o", str(exe), "-lrt"])
    key = "%x" % (0xb000 + os.getpid() % 4096)
    subprocess.check_call([_need("ringdb"), "-k", key, "-n", "4", "-b", "1048576"])
    try:
        w = subprocess.Popen([_need("junkdb"), "-k", key, "-r", "0", "-n", "11"], stdout=subprocess.PIPE, text=True)
        r = subprocess.run([str(exe), key], capture_output=True, text=True, timeout=60)
        w.wait(timeout=60)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "blocks 11 bufsz 1048576" in r.stdout
    finally:
        subprocess.call([_need("ringdb"), "-d", "-k", key])


@pytest.mark.gpu
def test_debug_program_matches_oracle_and_reference_binary(oracle, config_dir, tmp_path):
    """Config 1 end to end: bin/beam_debug on the linear array with 1024 sources writes bin/data.py
    [source][beam] (DM-0 sums).  It must equal the oracle's float path to the 6 digits the file
    holds, and the reference's own DEBUG binary rebuilt for sm_100 (oracle/_ref/beam_ref), run here on
    the same files."""
    cfg = [os.path.join(config_dir, f) for f in ("linear_positions.txt", "linear_directions.txt", "linear_source_directions_1024.txt")]
    work = tmp_path / "ours"; (work / "bin").mkdir(parents=True)
    r = subprocess.run([_need("beam_debug"), "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=work, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for milestone in ("Generating new source data", "obs Complete", "Observation ran in", "Time per data chunk", "Approximate datarate",
                      "Synchronized", "Freed GPU memory", "Freed CPU memory"):
        assert milestone in r.stdout, milestone
    ours = _load_data_py(work / "bin" / "data.py")
    assert ours.shape == (1024, 256)
    pos = oracle.read_positions(cfg[0]); dirs = oracle.read_directions(cfg[1], 256); src = oracle.read_directions(cfg[2])
    w = oracle.weights(pos, dirs, 0)
    data = oracle.generate(pos, src)
    for s in (0, 1, 255, 513, 777, 1023):
        ref = oracle.dm0(oracle.detect_ref_float(oracle.beamform(data[s], w), 2))
        assert np.allclose(ours[s], ref, rtol=2e-5), s          # 6 significant digits in the file + fp32 sum order
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "beam_ref")
    if os.path.exists(ref_exe):
        rw = tmp_path / "ref"; (rw / "bin").mkdir(parents=True)
        rr = subprocess.run([ref_exe, "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=rw, capture_output=True, text=True, timeout=900)
        assert rr.returncode == 0, rr.stdout[-1500:] + rr.stderr[-1500:]
        theirs = _load_data_py(rw / "bin" / "data.py")
        assert theirs.shape == ours.shape
        assert np.allclose(ours, theirs, rtol=2e-5)
        m = re.search(r"Time per data chunk: ([0-9.e+-]+)", rr.stdout)
        print("reference (sm_100 rebuild) time per data chunk [ms]:", m.group(1) if m else "?")


@pytest.mark.gpu
def test_debug_program_scaled_array(oracle, config_dir, tmp_path):
    """BASELINE.json configs[3] through the drop-in program: 256 antennas x 2048 beams (16 channels of the band),
    64 seeded sources; bin/data.py [source][beam] against the oracle's float path."""
    cfg = [os.path.join(config_dir, f) for f in ("scaled_positions_256.txt", "scaled_beam_directions_2048.txt", "scaled_source_directions_64.txt")]
    work = tmp_path / "scaled"; (work / "bin").mkdir(parents=True)
    env = dict(os.environ, DSABF_N_ANTENNAS="256", DSABF_N_BEAMS="2048", DSABF_N_CHANNELS="16")
    r = subprocess.run([_need("beam_debug"), "-g", "3", "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=work, capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ours = _load_data_py(work / "bin" / "data.py")
    assert ours.shape == (64, 2048)
    pos = oracle.read_positions(cfg[0], 256); dirs = oracle.read_directions(cfg[1], 2048); src = oracle.read_directions(cfg[2])
    w = oracle.weights(pos, dirs, 3, n_chan=16)
    data = oracle.generate(pos, src, n_chunks=64, n_chan=16, gpu=3)
    for s in (0, 1, 31, 63):
        ref = oracle.dm0(oracle.detect_ref_float(oracle.beamform(data[s], w), 2))
        assert np.allclose(ours[s], ref, rtol=2e-5), s
        assert ref.max() > 50 * np.median(ref)                  # a point source lights up a beam


@pytest.mark.gpu
def test_observation_program_streams_from_ring(config_dir):
    """Config 3: bin/beam reads 134217728-byte blocks from the ring (25-block burn-in, then data),
    copies them through pinned memory into the device ring, beamforms, and ends at the short block."""
    key = "%x" % (0xc000 + os.getpid() % 4096)
    subprocess.check_call([_need("ringdb"), "-k", key, "-n", "8", "-b", "134217728"])
    try:
        w = subprocess.Popen([_need("junkdb"), "-k", key, "-r", "0", "-n", "40", os.path.join(config_dir, "correlator_header_dsaX.txt")],
                             stdout=subprocess.PIPE, text=True)
        env = dict(os.environ, DSABF_REPORT="1")
        r = subprocess.run([_need("beam"), "-k", key, "-c", "-1", "-g", "2", "-p", os.path.join(config_dir, "grid_positions.txt"),
                            "-d", os.path.join(config_dir, "grid_beam_directions.txt")], capture_output=True, text=True, timeout=600, env=env)
        w.wait(timeout=60)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "block size is: 134217728" in r.stdout and "Done Burn in" in r.stdout and "obs Complete" in r.stdout
        m = re.search(r"Code produced outputs for (\d+) data chunks", r.stdout)
        assert m and int(m.group(1)) == 15 * 32 * 8          # 40 blocks - 25 burn-in = 15 blocks x 32 chunks x 8 outputs
    finally:
        subprocess.call([_need("ringdb"), "-d", "-k", key])
