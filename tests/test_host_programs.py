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


_READER_SRC = r"""
// reads a ring through the PSRDADA API (host/psrdada_compat + libpsrdada_shim.a), like the reference's dada_handler
#include "dada_hdu.h"
#include <cstdio>
int main(int argc, char** argv) {
    unsigned key; sscanf(argv[1], "%%x", &key);
    multilog_t* log = multilog_open("reader", 0);
    dada_hdu_t* h = dada_hdu_create(log);
    dada_hdu_set_key(h, (key_t)key);
    if (dada_hdu_connect(h) < 0 || dada_hdu_lock_read(h) < 0) return 2;
    uint64_t hs = 0; char* hdr = ipcbuf_get_next_read(h->header_block, &hs);
    if (!hdr) return 4;
    printf("header: %%.200s\n", hdr);
    ipcbuf_mark_cleared(h->header_block);
    uint64_t n = 0, bytes = 0, id = 0; unsigned long long sum = 0;
    for (;;) {
        char* b = ipcio_open_block_read(h->data_block, &bytes, &id);
        if (!b) break;
        if (bytes == 0) { ipcio_close_block_read(h->data_block, 0); break; }
        if (id != n) return 3;
        if (ipcio_open_block_read(h->data_block, &bytes, &id) != 0) return 5;     // PSRDADA: one open block per reader
        %s
        n++; ipcio_close_block_read(h->data_block, bytes);
    }
    printf("blocks %%llu bufsz %%llu sum %%llu\n", (unsigned long long)n, (unsigned long long)ipcbuf_get_bufsz((ipcbuf_t*)h->data_block), sum);
    dada_hdu_unlock_read(h); dada_hdu_destroy(h); return 0; }
"""


def _build_ring_reader(tmp_path, name, per_block):
    host = os.path.join(ROOT, "dsabeamformer_b200", "host")
    src = tmp_path / (name + ".cc")
    src.write_text(_READER_SRC % per_block)
    exe = tmp_path / name
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(host, "psrdada_compat"), str(src), "-o", str(exe),
                           os.path.join(host, "libpsrdada_shim.a"), "-lrt", "-lpthread"])
    return exe


def test_ring_tools_roundtrip(tmp_path):
    """ringdb + junkdb (stand-ins for dada_db / dada_junkdb): blocks arrive in order, full size, then EOD."""
    exe = _build_ring_reader(tmp_path, "reader", """
        for (uint64_t i = 0; i < bytes; i += 4096) sum += (unsigned char)b[i];
""")
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


def test_observation_loop_state_machine(tmp_path):
    """observation_loop.hh against stubbed tickets (no GPU): transfers are queued only while TQ-A < 4 and TQ-T < 2
    (MAX_TOTAL_SEP / MAX_TRANSFER_SEP, beamformer.hh:117-118), analyses only while AQ < T, work retires in ticket
    order, ring slots are the counters modulo 8, and the DEBUG run completes once every source had its chunk."""
    src = tmp_path / "loop.cc"
    src.write_text('''
#include "observation_loop.hh"
#include <cstdio>
#include <cstdlib>
static uint64_t done_t = 0, done_a = 0;                 // tickets below these values have completed
extern "C" int dsabf_query_transfer(dsabf_handle*, uint64_t t) { return t < done_t ? 1 : 0; }
extern "C" int dsabf_query_analysis(dsabf_handle*, uint64_t t) { return t < done_a ? 1 : 0; }
#define REQUIRE(c) do { if (!(c)) { printf("FAILED line %d: %s\\n", __LINE__, #c); return 1; } } while (0)
int main() {
    beam::ObservationLoopState st(nullptr, 2, 4, 32, 8, true);
    st.set_n_pt_sources(10 * 32);                         // 10 blocks of sources
    uint64_t next_t = 0, next_a = 0; unsigned rng = 12345; int steps = 0;
    while (!st.observations_complete()) {
        REQUIRE(++steps < 100000);
        rng = rng * 1664525u + 1013904223u;
        if (st.ready_for_transfer() && !st.check_transfers_complete_debug()) {
            REQUIRE(st.blocks_transfer_queue() - st.blocks_analyzed() < 4);
            REQUIRE(st.blocks_transfer_queue() - st.blocks_transferred() < 2);
            REQUIRE(st.next_transfer_slot() == (int)(st.blocks_transfer_queue() % 8));
            st.transfer_enqueued(next_t++);
        }
        if (st.ready_for_analysis()) {
            REQUIRE(st.blocks_analysis_queue() < st.blocks_transferred());
            REQUIRE(st.next_analysis_slot() == (int)(st.blocks_analysis_queue() % 8));
            REQUIRE(st.current_analysis_chunk(31) == st.blocks_analysis_queue() * 32 + 31);
            st.analysis_enqueued(next_a++);
        }
        if ((rng >> 16) % 3 == 0 && done_t < next_t) done_t++;       // copies and kernels finish at random times, in order
        if ((rng >> 20) % 4 == 0 && done_a < next_a) done_a++;
        const uint64_t t0 = st.blocks_transferred(), a0 = st.blocks_analyzed();
        st.check_transfer_events(false); st.check_analysis_events(false);
        REQUIRE(st.blocks_transferred() == done_t && st.blocks_analyzed() == done_a);
        REQUIRE(st.blocks_transferred() >= t0 && st.blocks_analyzed() >= a0);
        REQUIRE(st.blocks_analyzed() <= st.blocks_analysis_queue() && st.blocks_analysis_queue() <= st.blocks_transferred());
        REQUIRE(st.blocks_transferred() <= st.blocks_transfer_queue());
    }
    REQUIRE(st.blocks_analyzed() == 10 && st.blocks_transfer_queue() == 10 && st.transfers_complete());
    printf("ok %d steps\\n", steps);
    return 0;
}
''')
    exe = tmp_path / "loop"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "dsabeamformer_b200", "host"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


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
        # The reference's OWN main() with the INTEGRATION.md drop-in applied (oracle/_ref/beam_ref_dsabf: its option
        # parsing, generator, loop state, H2D ring, streams, gemv and data.py writer; expand + cuBLAS GEMM + detect replaced
        # by dsabf_beamform_device on the same stream): same file, to the 6 digits it holds.
        patched = os.path.join(ROOT, "oracle", "_ref", "beam_ref_dsabf")
        if os.path.exists(patched):
            pw = tmp_path / "patched"; (pw / "bin").mkdir(parents=True)
            pr = subprocess.run([patched, "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=pw, capture_output=True, text=True, timeout=900)
            assert pr.returncode == 0, pr.stdout[-1500:] + pr.stderr[-1500:]
            mine = _load_data_py(pw / "bin" / "data.py")
            assert mine.shape == theirs.shape and np.allclose(mine, theirs, rtol=2e-5)
            assert np.allclose(mine, ours, rtol=2e-5)          # same fused kernel as bin/beam_debug (chunk by chunk); DM-0 sum by cuBLAS gemv there, dm0_kernel here
            m = re.search(r"Time per data chunk: ([0-9.e+-]+)", pr.stdout)
            print("reference main() + libdsabf time per data chunk [ms]:", m.group(1) if m else "?")


@pytest.mark.gpu
def test_debug_program_grid_4096_sources(oracle, config_dir, tmp_path):
    """Config 2 end to end (BASELINE.json configs[1]): the 2-D grid array with 4096 source directions drives the
    regenerate-in-place path four times (src/beamformer.cu:405-418, test_data_generator.hh:98-108): 4 generator batches
    = 128 blocks.  bin/data.py [4096][256] against the oracle (sources from every batch) and against the reference's own
    DEBUG binary on the same files; once more with the CPU generator (DSABF_CPU_GENERATOR=1): identical file."""
    cfg = [os.path.join(config_dir, f) for f in ("grid_positions.txt", "grid_beam_directions.txt", "grid_source_directions_4096.txt")]
    work = tmp_path / "ours"; (work / "bin").mkdir(parents=True)
    r = subprocess.run([_need("beam_debug"), "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=work, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("Generating new source data") == 4 and "obs Complete" in r.stdout
    m = re.search(r"Code produced outputs for (\d+) data chunks", r.stdout)
    assert m and int(m.group(1)) == 4096 * 8
    t_ours = float(re.search(r"Observation ran in ([0-9.e+-]+)", r.stdout).group(1))
    ours = _load_data_py(work / "bin" / "data.py")
    assert ours.shape == (4096, 256)
    pos = oracle.read_positions(cfg[0]); dirs = oracle.read_directions(cfg[1], 256); src = oracle.read_directions(cfg[2])
    w = oracle.weights(pos, dirs, 0)
    for b in range(4):
        data = oracle.generate(pos, src, batch=b)
        for s in (0, 517, 1023):
            ref = oracle.dm0(oracle.detect_ref_float(oracle.beamform(data[s], w), 2))
            assert np.allclose(ours[1024 * b + s], ref, rtol=2e-5), (b, s)
    cw = tmp_path / "cpu"; (cw / "bin").mkdir(parents=True)
    rc = subprocess.run([_need("beam_debug"), "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=cw, capture_output=True, text=True, timeout=1800,
                        env=dict(os.environ, DSABF_CPU_GENERATOR="1"))
    assert rc.returncode == 0, rc.stdout[-2000:] + rc.stderr[-2000:]
    assert (cw / "bin" / "data.py").read_bytes() == (work / "bin" / "data.py").read_bytes()
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "beam_ref")
    if os.path.exists(ref_exe):
        rw = tmp_path / "ref"; (rw / "bin").mkdir(parents=True)
        rr = subprocess.run([ref_exe, "-p", cfg[0], "-d", cfg[1], "-s", cfg[2]], cwd=rw, capture_output=True, text=True, timeout=3000)
        assert rr.returncode == 0, rr.stdout[-1500:] + rr.stderr[-1500:]
        theirs = _load_data_py(rw / "bin" / "data.py")
        assert theirs.shape == ours.shape and np.allclose(ours, theirs, rtol=2e-5)
        t_ref = float(re.search(r"Observation ran in ([0-9.e+-]+)", rr.stdout).group(1))
        print("config 2 (4096 grid sources): observation time ours %.1f ms, reference (sm_100 rebuild) %.1f ms" % (t_ours, t_ref))


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


@pytest.mark.gpu
def test_reference_observation_program_reads_our_ring(config_dir):
    """Link- and run-compatibility of the PSRDADA shim: the reference's own observation build (src/beamformer.cu with
    its dada_handler.hh, compiled unmodified against host/psrdada_compat + libpsrdada_shim by oracle/Makefile) attaches
    to a ring made by bin/ringdb, pins it, burns in 25 blocks, consumes the data blocks bin/junkdb writes and stops at
    the end-of-data block -- the same ring, header and blocks bin/beam reads."""
    exe = os.path.join(ROOT, "oracle", "_ref", "beam_ref_obs")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/beam_ref_obs not built (needs /root/reference at build time)")
    key = "%x" % (0xa000 + os.getpid() % 4096)
    subprocess.check_call([_need("ringdb"), "-k", key, "-n", "8", "-b", "134217728"])
    try:
        w = subprocess.Popen([_need("junkdb"), "-k", key, "-r", "0", "-n", "29", os.path.join(config_dir, "correlator_header_dsaX.txt")],
                             stdout=subprocess.PIPE, text=True)
        r = subprocess.run([exe, "-k", key, "-c", "-1", "-g", "0"], capture_output=True, text=True, timeout=600)
        wout, _ = w.communicate(timeout=60)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "block size is: 134217728" in r.stdout and "Done Burn in" in r.stdout and "Freed CPU memory" in r.stdout
        assert "wrote 29 blocks" in wout              # the writer only gets there if the reader released every block
    finally:
        subprocess.call([_need("ringdb"), "-d", "-k", key])


def _splitmix64_block(n_bytes, seed=0xD5A):
    """the junk block of bin/junkdb (ring_tools.cc): splitmix64 stream, little-endian, vectorised"""
    with np.errstate(over="ignore"):
        i = np.arange(1, n_bytes // 8 + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)


@pytest.mark.gpu
def test_observation_program_output_sink(bf, oracle, config_dir, tmp_path):
    """with DSABF_OUT_KEY bin/beam appends every finished block of power (float [time][chan][beam],
    64 MiB) to an output ring.  Three data blocks after the burn-in; the blocks read back must equal what the library
    computes for the same ring input (junkdb's splitmix64 block, rotated by (block % 16) * 4096 bytes)."""
    exe = _build_ring_reader(tmp_path, "dump", """
        { char name[256]; snprintf(name, sizeof name, "%s/block%llu.bin", argv[2], (unsigned long long)n);
          FILE* f = fopen(name, "wb"); fwrite(b, 1, bytes, f); fclose(f); }
""")
    key_in, key_out = "%x" % (0xd000 + os.getpid() % 4096), "%x" % (0xe000 + os.getpid() % 4096)
    subprocess.check_call([_need("ringdb"), "-k", key_in, "-n", "8", "-b", "134217728"])
    subprocess.check_call([_need("ringdb"), "-k", key_out, "-n", "4", "-b", "67108864"])
    try:
        w = subprocess.Popen([_need("junkdb"), "-k", key_in, "-r", "0", "-n", "28", os.path.join(config_dir, "correlator_header_dsaX.txt")],
                             stdout=subprocess.PIPE, text=True)
        pfile, dfile = os.path.join(config_dir, "grid_positions.txt"), os.path.join(config_dir, "grid_beam_directions.txt")
        r = subprocess.run([_need("beam"), "-k", key_in, "-c", "-1", "-g", "2", "-p", pfile, "-d", dfile], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, DSABF_OUT_KEY=key_out))
        w.wait(timeout=60)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        d = subprocess.run([str(exe), key_out, str(tmp_path)], capture_output=True, text=True, timeout=120)
        assert d.returncode == 0 and "blocks 3" in d.stdout, d.stdout + d.stderr
        hdr = dict(l.split(None, 1) for l in d.stdout.splitlines() if len(l.split(None, 1)) == 2)
        assert hdr.get("ORDER") == "TFB" and hdr.get("NBEAM") == "256" and hdr.get("SUBBAND") == "2" and hdr.get("NTIME") == "256", d.stdout
    finally:
        subprocess.call([_need("ringdb"), "-d", "-k", key_in])
        subprocess.call([_need("ringdb"), "-d", "-k", key_out])
    junk = _splitmix64_block(134217728)
    pos = oracle.read_positions(pfile); dirs = oracle.read_directions(dfile, 256)
    B = bf.Beamformer(debug=False)
    w = bf.compute_weights(pos, dirs, 2)
    B.set_weights(w)
    for i, blk in ((0, 25), (2, 27)):                      # first and last data block (ring blocks 25 and 27 after the burn-in)
        packed = np.roll(junk, -((blk % 16) * 4096)).reshape(32, 256, 256, 64)
        got = np.fromfile(tmp_path / ("block%d.bin" % i), dtype=np.float32).reshape(256, 256, 256)
        assert np.array_equal(got, B.beamform(packed)), i
        # and against the oracle (the reference's float detect_sum on exact int32 voltages) for two chunks of the block
        for c in (0, 31):
            ref = oracle.detect_ref_float(oracle.beamform(packed[c], w), 32)
            rel = np.abs(got[c * 8:(c + 1) * 8] - ref) / np.maximum(ref, 1e-30)
            assert rel.max() <= 1e-6, (i, c, rel.max())
    B.close()

