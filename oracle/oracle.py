"""ctypes binding of oracle/_build/liboracle.so + config-file parsing.

TEST INFRASTRUCTURE -- only tests/, bench.py's cpu_baseline / --impl reference
legs and __graft_entry__.smoke() import this module.  The product package
(dsabeamformer_b200) never does.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

N_ANT, N_BEAMS, N_CHAN = 64, 256, 256


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "dsabf_oracle.c")):
        build()
    lib = ctypes.CDLL(_LIB)
    lib.orc_freq_weights.restype = ctypes.c_float
    lib.orc_freq_generator.restype = ctypes.c_float
    lib.orc_dm_trials.restype = ctypes.c_int
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a, t=ctypes.c_void_p):
    return a.ctypes.data_as(t)


# --- config files: /root/reference/src/beamformer.hh:250-284 (operator>> into float) ---
def read_table(path, ncol, expected=None):
    """count, then `ncol` floats per row; rows beyond the file stay 0; parsing
    stops at the first token that is not a number (as istream >> float does)."""
    toks = open(path, encoding="utf-8", errors="replace").read().split()
    n = int(toks[0])
    rows = n if expected is None else expected
    out = np.zeros((rows, ncol), dtype=np.float32)
    vals = []
    for t in toks[1:]:
        try:
            vals.append(np.float32(float(t)))
        except ValueError:
            break
    vals = np.array(vals[: rows * ncol], dtype=np.float32)
    out.reshape(-1)[: len(vals)] = vals
    return n, out


def read_positions(path, n_ant=N_ANT):
    return read_table(path, 3, n_ant)[1]


def read_directions(path, n=None):
    return read_table(path, 2, n)[1]


def default_positions(n_ant=N_ANT):
    """src/beamformer.cu:135-140: x = i*500.0/(N-1) - 250.0 (double) stored to float."""
    pos = np.zeros((n_ant, 3), dtype=np.float32)
    pos[:, 0] = np.array([i * 500.0 / (n_ant - 1) - 250.0 for i in range(n_ant)], dtype=np.float64).astype(np.float32)
    return pos


def default_directions(n_beams=N_BEAMS):
    """src/beamformer.cu:142-147 with PI = 3.14159265358979, HALF_FOV = 3.5."""
    PI = 3.14159265358979
    d2r = lambda x: x * PI / 180.0
    d = np.zeros((n_beams, 2), dtype=np.float32)
    d[:, 0] = np.array([i * d2r(2 * 3.5) / (n_beams - 1) - d2r(3.5) for i in range(n_beams)], dtype=np.float64).astype(np.float32)
    return d


def freq_weights(gpu, n_chan=N_CHAN):
    return np.array([lib().orc_freq_weights(gpu, i) for i in range(n_chan)], dtype=np.float32)


def freq_generator(gpu, n_chan=N_CHAN):
    return np.array([lib().orc_freq_generator(gpu, i) for i in range(n_chan)], dtype=np.float32)


def weights(pos, dirs, gpu=0, n_chan=N_CHAN):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n_ant, n_beams = pos.shape[0], dirs.shape[0]
    w = np.empty((n_chan, n_ant, n_beams, 2), dtype=np.int8)
    lib().orc_weights(_p(pos), _p(dirs), n_ant, n_beams, n_chan, gpu, _p(w))
    return w


def generate(pos, src, batch=0, n_chunks=1024, n_chan=N_CHAN, cols=16, gpu=0):
    """data[n_chunks][n_chan][cols][n_ant] uint8; src=None -> bogus 0x70 fill."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n_ant = pos.shape[0]
    data = np.empty((n_chunks, n_chan, cols, n_ant), dtype=np.uint8)
    if src is None:
        lib().orc_generate(_p(pos), None, 0, batch, n_chunks, n_ant, n_chan, cols, gpu, _p(data))
    else:
        src = np.ascontiguousarray(src, dtype=np.float32)
        lib().orc_generate(_p(pos), _p(src), src.shape[0], batch, n_chunks, n_ant, n_chan, cols, gpu, _p(data))
    return data


def unpack(packed):
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(packed.shape + (2,), dtype=np.int8)
    lib().orc_unpack(_p(packed), _p(out), ctypes.c_size_t(packed.size))
    return out


def beamform(packed, w):
    """packed [n_chan][n_cols][n_ant] u8, w [n_chan][n_ant][n_beams][2] i8 ->
    V [n_chan][n_cols][n_beams][2] int32."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    w = np.ascontiguousarray(w, dtype=np.int8)
    n_chan, n_cols, n_ant = packed.shape
    n_beams = w.shape[2]
    assert w.shape[0] == n_chan and w.shape[1] == n_ant
    V = np.empty((n_chan, n_cols, n_beams, 2), dtype=np.int32)
    lib().orc_beamform(_p(packed), _p(w), n_chan, n_cols, n_ant, n_beams, _p(V))
    return V


def power_int(V, n_avg):
    n_chan, n_cols, n_beams, _ = V.shape
    P = np.empty((n_cols // n_avg, n_chan, n_beams), dtype=np.int64)
    lib().orc_power_int(_p(np.ascontiguousarray(V)), n_chan, n_cols, n_beams, n_avg, _p(P))
    return P


def detect_ref_float(V, n_avg):
    n_chan, n_cols, n_beams, _ = V.shape
    out = np.empty((n_cols // n_avg, n_chan, n_beams), dtype=np.float32)
    lib().orc_detect_ref_float(_p(np.ascontiguousarray(V)), n_chan, n_cols, n_beams, n_avg, _p(out))
    return out


def dm0(out, o=0):
    n_out, n_chan, n_beams = out.shape
    ded = np.empty(n_beams, dtype=np.float32)
    lib().orc_dm0(_p(np.ascontiguousarray(out, dtype=np.float32)), n_chan, n_beams, o, _p(ded))
    return ded


def dedisperse(power, delay, n_t_out):
    power = np.ascontiguousarray(power, dtype=np.float32)
    delay = np.ascontiguousarray(delay, dtype=np.int32)
    n_t, n_chan, n_beams = power.shape
    n_dm = delay.shape[0]
    ded = np.empty((n_dm, n_t_out, n_beams), dtype=np.float32)
    lib().orc_dedisperse(_p(power), n_t, n_chan, n_beams, _p(delay), n_dm, n_t_out, _p(ded))
    return ded


def dm_trials(dm_max=2000.0, nchan=2048):
    buf = np.zeros(8192, dtype=np.float64)
    n = lib().orc_dm_trials(ctypes.c_double(dm_max), nchan, _p(buf), len(buf))
    return buf[:n].copy()


def dm_delays(dms, freq_ghz, f1_ghz=1.53, tsamp_ms_times_navg=0.131 * 16):
    dms = np.ascontiguousarray(dms, dtype=np.float64)
    freq_ghz = np.ascontiguousarray(freq_ghz, dtype=np.float64)
    delay = np.empty((len(dms), len(freq_ghz)), dtype=np.int32)
    lib().orc_dm_delays(_p(dms), len(dms), _p(freq_ghz), len(freq_ghz), ctypes.c_double(f1_ghz),
                        ctypes.c_double(tsamp_ms_times_navg), _p(delay))
    return delay


def chunk_pipeline(chunk, w, n_avg):
    """One chunk [n_chan][cols][n_ant] -> (V, P_int, out_ref_float)."""
    V = beamform(chunk, w)
    return V, power_int(V, n_avg), detect_ref_float(V, n_avg)
