"""dsabeamformer_b200 -- Python mirror of the C ABI in include/dsabf.h.

The product is the sm_100a CUDA library `csrc/libdsabf.so`; this module is only
a ctypes binding so that tests and bench.py can drive exactly the entry points
the reference's C++ observation loop would call
(/root/reference/src/beamformer.cu:461-521 and the set-up around it).  There is
NO CPU fallback: importing works anywhere, but every compute call fails loudly
if the library or a B200-class GPU is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("DSABF_LIB") or os.path.join(CSRC, "libdsabf.so")    # DSABF_LIB: developer hook for instrumented builds

E_INVALID, E_STATE, E_WEIGHT = 10001, 10002, 10003


class DsabfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dsabf error %d: %s" % (code, msg))
        self.code = code


class Config(ctypes.Structure):
    """struct dsabf_config (shape constants of src/beamformer.hh:47-60,108-147)."""
    _fields_ = [
        ("n_antennas", ctypes.c_int), ("n_beams", ctypes.c_int), ("n_channels", ctypes.c_int),
        ("n_pol", ctypes.c_int), ("n_averaging", ctypes.c_int), ("n_outputs_per_chunk", ctypes.c_int),
        ("n_chunks_per_block", ctypes.c_int), ("n_ring_slots", ctypes.c_int), ("device", ctypes.c_int),
        ("exact_power", ctypes.c_int), ("power_scale", ctypes.c_float),
    ]


def build(verbose=False):
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]) + ["-j4"])


_lib = None


def lib():
    """Load libdsabf.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DsabfError(-1, "%s is missing: run `make -C %s` (or __graft_entry__.build())" % (LIB_PATH, CSRC))
    L = ctypes.CDLL(LIB_PATH)
    vp, i, u64p = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)
    cfgp = ctypes.POINTER(Config)
    L.dsabf_default_config.argtypes = [cfgp, i]
    L.dsabf_chunk_bytes.argtypes = [cfgp]; L.dsabf_chunk_bytes.restype = ctypes.c_size_t
    L.dsabf_block_bytes.argtypes = [cfgp]; L.dsabf_block_bytes.restype = ctypes.c_size_t
    L.dsabf_chunk_out_floats.argtypes = [cfgp]; L.dsabf_chunk_out_floats.restype = ctypes.c_size_t
    L.dsabf_columns_per_chunk.argtypes = [cfgp]
    L.dsabf_create.argtypes = [cfgp, ctypes.POINTER(vp)]
    L.dsabf_destroy.argtypes = [vp]
    L.dsabf_last_error_string.restype = ctypes.c_char_p
    L.dsabf_compute_weights.argtypes = [vp, vp, i, i, i, i, vp]
    L.dsabf_set_weights.argtypes = [vp, vp]
    L.dsabf_ring_slot_ptr.argtypes = [vp, i]; L.dsabf_ring_slot_ptr.restype = vp
    L.dsabf_enqueue_h2d.argtypes = [vp, i, vp, ctypes.c_size_t, u64p]
    L.dsabf_query_transfer.argtypes = [vp, ctypes.c_uint64]
    L.dsabf_query_analysis.argtypes = [vp, ctypes.c_uint64]
    L.dsabf_enqueue_block.argtypes = [vp, i, i, vp, vp, i, u64p]
    L.dsabf_block_out_ptr.argtypes = [vp, i]; L.dsabf_block_out_ptr.restype = vp
    L.dsabf_sync.argtypes = [vp]
    L.dsabf_beamform_device.argtypes = [vp, vp, i, vp, vp]
    L.dsabf_beamform_host.argtypes = [vp, vp, i, vp]
    L.dsabf_debug_voltages_host.argtypes = [vp, vp, i, vp, vp]
    L.dsabf_dedisperse_device.argtypes = [vp, vp, i, i, i, vp, i, i, vp, vp]
    L.dsabf_transpose_pull_device.argtypes = [vp, vp, i, i, i, i, i, vp, vp]
    L.dsabf_dedisperse_single_peers_device.argtypes = [vp, vp, i, i, i, i, i, vp, i, vp, vp]
    L.dsabf_dedisperse_host.argtypes = [vp, vp, i, i, i, vp, i, i, vp]
    L.dsabf_generate_test_data_device.argtypes = [vp, vp, vp, i, i, i, vp, vp]
    L.dsabf_generate_test_data_host.argtypes = [vp, vp, vp, i, i, i, vp]
    L.dsabf_dm_trials.argtypes = [ctypes.c_double, i, vp, i]
    L.dsabf_dm_delays.argtypes = [vp, i, vp, i, vp]
    L.dsabf_kernel_launches.argtypes = [vp]; L.dsabf_kernel_launches.restype = ctypes.c_uint64
    L.dsabf_fused_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i), i]
    L.dsabf_enable_kernel_timing.argtypes = [vp, i]
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise DsabfError(rc, lib().dsabf_last_error_string().decode())


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def default_config(debug=False, **overrides):
    cfg = Config()
    _chk(lib().dsabf_default_config(ctypes.byref(cfg), 1 if debug else 0))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def compute_weights(pos, dirs, gpu=0, n_chan=256):
    """Steering weights, the host loop of src/beamformer.cu:230-241.
    pos [n_ant][3] f32, dirs [n_beams][2] f32 -> int8 [n_chan][n_ant][n_beams][2]."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    w = np.empty((n_chan, pos.shape[0], dirs.shape[0], 2), dtype=np.int8)
    _chk(lib().dsabf_compute_weights(_ptr(pos), _ptr(dirs), pos.shape[0], dirs.shape[0], n_chan, gpu, _ptr(w)))
    return w


def dm_trials(dm_max=2000.0, n_chan_total=2048):
    """DM trial list of sandbox/Dispersion Theory.ipynb cells 1-2 (1627 trials to DM 2000 at 2048 channels)."""
    n = lib().dsabf_dm_trials(dm_max, n_chan_total, None, 0)
    if n < 0:
        raise DsabfError(n, lib().dsabf_last_error_string().decode())
    dms = np.empty(n, dtype=np.float64)
    lib().dsabf_dm_trials(dm_max, n_chan_total, _ptr(dms), n)
    return dms


def dm_delays(dms, freq_ghz):
    """int32 [n_dm][n_chan] delays in output samples (cell 5 of the same notebook)."""
    dms = np.ascontiguousarray(dms, dtype=np.float64)
    freq_ghz = np.ascontiguousarray(freq_ghz, dtype=np.float64)
    delay = np.empty((len(dms), len(freq_ghz)), dtype=np.int32)
    _chk(lib().dsabf_dm_delays(_ptr(dms), len(dms), _ptr(freq_ghz), len(freq_ghz), _ptr(delay)))
    return delay


class Beamformer:
    """One handle == one GPU == the reference's one `beam` process."""

    def __init__(self, debug=False, **overrides):
        self.cfg = default_config(debug, **overrides)
        self._h = ctypes.c_void_p()
        _chk(lib().dsabf_create(ctypes.byref(self.cfg), ctypes.byref(self._h)))
        c = self.cfg
        self.cols_per_chunk = lib().dsabf_columns_per_chunk(ctypes.byref(c))
        self.chunk_bytes = lib().dsabf_chunk_bytes(ctypes.byref(c))
        self.block_bytes = lib().dsabf_block_bytes(ctypes.byref(c))
        self.chunk_out_floats = lib().dsabf_chunk_out_floats(ctypes.byref(c))
        self.group = c.n_pol * c.n_averaging

    def close(self):
        if self._h:
            lib().dsabf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights --
    def set_weights(self, w):
        c = self.cfg
        w = np.ascontiguousarray(w, dtype=np.int8)
        assert w.shape == (c.n_channels, c.n_antennas, c.n_beams, 2), w.shape
        _chk(lib().dsabf_set_weights(self._h, _ptr(w)))

    # -- synchronous host-buffer path --
    def _as_chunks(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        assert packed.size % self.chunk_bytes == 0, (packed.size, self.chunk_bytes)
        return packed, packed.size // self.chunk_bytes

    def beamform(self, packed):
        """packed [n_chunks][chan][col][ant] u8 -> power [n_chunks*n_out][chan][beam] f32."""
        packed, n = self._as_chunks(packed)
        c = self.cfg
        out = np.empty((n * c.n_outputs_per_chunk, c.n_channels, c.n_beams), dtype=np.float32)
        _chk(lib().dsabf_beamform_host(self._h, _ptr(packed), n, _ptr(out)))
        return out

    def voltages(self, packed, want_power=True):
        """-> V int32 [chan][n_chunks*cols][beam][2], P int64 [n_chunks*n_out][chan][beam]."""
        packed, n = self._as_chunks(packed)
        c = self.cfg
        V = np.empty((c.n_channels, n * self.cols_per_chunk, c.n_beams, 2), dtype=np.int32)
        P = np.empty((n * c.n_outputs_per_chunk, c.n_channels, c.n_beams), dtype=np.int64) if want_power else None
        _chk(lib().dsabf_debug_voltages_host(self._h, _ptr(packed), n, _ptr(V), _ptr(P) if want_power else None))
        return V, P

    # -- streaming path (the observation loop's calls) --
    def enqueue_h2d(self, slot, host_ptr, nbytes):
        t = ctypes.c_uint64()
        _chk(lib().dsabf_enqueue_h2d(self._h, slot, ctypes.c_void_p(host_ptr), nbytes, ctypes.byref(t)))
        return t.value

    def enqueue_block(self, slot, n_chunks, beam_out_ptr=None, ded_ptr=None, n_valid_dedisp=0):
        t = ctypes.c_uint64()
        _chk(lib().dsabf_enqueue_block(self._h, slot, n_chunks, ctypes.c_void_p(beam_out_ptr), ctypes.c_void_p(ded_ptr),
                                       n_valid_dedisp, ctypes.byref(t)))
        return t.value

    def query_transfer(self, ticket):
        return lib().dsabf_query_transfer(self._h, ticket)

    def query_analysis(self, ticket):
        return lib().dsabf_query_analysis(self._h, ticket)

    def ring_slot_ptr(self, slot):
        return lib().dsabf_ring_slot_ptr(self._h, slot)

    def block_out_ptr(self, slot):
        return lib().dsabf_block_out_ptr(self._h, slot)

    def sync(self):
        _chk(lib().dsabf_sync(self._h))

    def beamform_device(self, d_packed_ptr, n_chunks, d_out_ptr, stream=None):
        _chk(lib().dsabf_beamform_device(self._h, ctypes.c_void_p(d_packed_ptr), n_chunks, ctypes.c_void_p(d_out_ptr),
                                         ctypes.c_void_p(stream)))

    # -- DEBUG-mode synthetic input (test_data_generator.hh:63-95) on the GPU --
    def generate_test_data(self, pos, sources, batch=0, gpu=0):
        """-> uint8 [1024][chan][col][ant]: batch `batch` of the point-source catalogue `sources` [n][2] (theta, phi)."""
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        sources = np.ascontiguousarray(sources, dtype=np.float32).reshape(-1, 2)
        c = self.cfg
        out = np.empty((1024, c.n_channels, self.cols_per_chunk, c.n_antennas), dtype=np.uint8)
        _chk(lib().dsabf_generate_test_data_host(self._h, _ptr(pos), _ptr(sources), sources.shape[0], batch, gpu, _ptr(out)))
        return out

    def generate_test_data_device(self, pos, sources, d_dst_ptr, batch=0, gpu=0, stream=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        sources = np.ascontiguousarray(sources, dtype=np.float32).reshape(-1, 2)
        _chk(lib().dsabf_generate_test_data_device(self._h, _ptr(pos), _ptr(sources), sources.shape[0], batch, gpu, ctypes.c_void_p(d_dst_ptr),
                                                   ctypes.c_void_p(stream)))

    # -- dedispersion --
    def dedisperse(self, power, delay, n_t_out):
        power = np.ascontiguousarray(power, dtype=np.float32)
        delay = np.ascontiguousarray(delay, dtype=np.int32)
        n_t, n_chan, n_beams = power.shape
        n_dm = delay.shape[0]
        assert delay.shape == (n_dm, n_chan)
        out = np.empty((n_dm, n_t_out, n_beams), dtype=np.float32)
        _chk(lib().dsabf_dedisperse_host(self._h, _ptr(power), n_t, n_chan, n_beams, _ptr(delay), n_dm, n_t_out, _ptr(out)))
        return out

    def dedisperse_device(self, d_power_ptr, n_t, n_chan, n_beams, delay, n_t_out, d_out_ptr, stream=None):
        delay = np.ascontiguousarray(delay, dtype=np.int32)
        _chk(lib().dsabf_dedisperse_device(self._h, ctypes.c_void_p(d_power_ptr), n_t, n_chan, n_beams, _ptr(delay),
                                           delay.shape[0], n_t_out, ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream)))

    def transpose_pull(self, peer_ptrs, rank, n_t, c_local, n_beams, d_full_ptr, stream=None):
        """full[t][g*c_local+c][b] = power_g[t][c][rank*Bs+b], pulled from the peers' buffers over NVLink (exchange.cu)."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _chk(lib().dsabf_transpose_pull_device(self._h, arr, len(peer_ptrs), rank, n_t, c_local, n_beams, ctypes.c_void_p(d_full_ptr),
                                               ctypes.c_void_p(stream)))

    def dedisperse_single_peers(self, peer_ptrs, rank, n_t, c_local, n_beams, delay, n_t_out, d_out_ptr, stream=None):
        """One trial over the full band, channels read from the peers' buffers in place (exchange fused into the sum)."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        delay = np.ascontiguousarray(delay, dtype=np.int32).reshape(-1)
        assert delay.shape[0] == len(peer_ptrs) * c_local
        _chk(lib().dsabf_dedisperse_single_peers_device(self._h, arr, len(peer_ptrs), rank, n_t, c_local, n_beams, _ptr(delay), n_t_out,
                                                        ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream)))

    # -- introspection --
    @property
    def kernel_launches(self):
        return int(lib().dsabf_kernel_launches(self._h))

    def enable_kernel_timing(self, on=True):
        _chk(lib().dsabf_enable_kernel_timing(self._h, 1 if on else 0))

    def fused_kernel_ms(self, reset=True):
        ms, n = ctypes.c_double(), ctypes.c_int()
        _chk(lib().dsabf_fused_kernel_ms(self._h, ctypes.byref(ms), ctypes.byref(n), 1 if reset else 0))
        return ms.value, n.value
