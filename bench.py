#!/usr/bin/env python3
"""bench.py -- beamformed samples/s (beams x chans x input-time) of the fused
unpack+beamform+detect hot path, per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sustain SECONDS] [--no-extras]

Workload (config.workload): the shape the reference's own number is quoted on
(README.md:11 -> BASELINE.md section 1) and BASELINE.json's configs[2]: 64 antennas
x 2 pol, 256 channels, 256 beams, N_AVERAGING = 16, PSRDADA-size blocks of
134 217 728 B (32 chunks), streamed.  One *step* = 8 such blocks (1 GiB of packed
voltages, > L2) through the hot path on every GPU; with N GPUs every rank owns one
256-channel sub-band (`-g rank`), no data-path collective (weak scaling).

  value : whole-job samples/s, inputs already resident in HBM, fused kernel only
          (dsabf_beamform_device), CUDA events on the launching stream, max over ranks (burst: K steps)
  value_sustained / sustained : the same loop kept running for --sustain seconds (default 5), with its own clock
          record and a roofline against the sustained peak
  e2e   : same metric through the C ABI with HOST buffers: pinned H2D of every block
          into the device ring, fused kernel, D2H of the float power into pinned memory
          (dsabf_enqueue_h2d / dsabf_enqueue_block), all inside the timed region; copy_ceiling is the bare
          pinned cudaMemcpyAsync H2D || D2H of the same bytes on the same ranks at the same time
  roofline : dominant kernel = fused_beamform_kernel; achieved = algorithmic int8 ops
          (8*N_ANT*N_POL = 1024 per sample) / mean kernel time (CUDA events on its stream);
          peak = the bare tcgen05.mma kind::i8 issue rate measured in this run (tools/ubench)
  cpu_baseline : the reference's numpy beamformer (oracle/numpy_beamformer.py, a port of
          sandbox/Beamformer Theory.ipynb) on this host's cores, bounded sample
  scaled_sharded : BASELINE.json configs[3], 256 ant x 2048 beams x 2048 channels split over the N ranks (strong scaling)
  dedisp_full_band : BASELINE.json configs[4], power [T][256 ch][256 beams] per rank -> frequency-to-beam exchange
          (one pull kernel over NVLink peer memory) -> 1627-trial sweep on 256/N beams per rank
--impl reference times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ANT, N_POL, N_CHAN, N_BEAMS, N_AVG = 64, 2, 256, 256, 16
CHUNKS_PER_BLOCK, BLOCKS_PER_STEP = 32, 8
OPS_PER_SAMPLE = 8 * N_ANT * N_POL
REALTIME_SAMPLES_PER_S = N_BEAMS * N_CHAN / 8.192e-6  # DSA real time per 256-channel GPU
WORKLOAD = ("PSRDADA-style streaming, observation shape: 64 ant x 2 pol x 256 chan -> 256 beams, "
            "N_AVERAGE=16, 134217728-byte blocks (BASELINE.json configs[2]; the shape README.md:11 is quoted on)")
METRIC = "beamformed samples/s (beams x chans x input-time)"
# one dict for both arms (the driver compares them)
CONFIG = {"workload": WORKLOAD, "n_antennas": N_ANT, "n_pol": N_POL, "n_channels_per_gpu": N_CHAN, "n_beams": N_BEAMS,
          "n_averaging": N_AVG, "blocks_per_step": BLOCKS_PER_STEP, "bytes_in_per_step_per_gpu": BLOCKS_PER_STEP * 134217728,
          "l2_policy": "inputs larger than L2 (1 GiB of packed voltages per step per GPU, 512 MiB of power out)",
          "parallelism": "channel sub-band per GPU, no collective", "power_mode": "fp32 (bit-exact below 2^24)"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------- CPU path
def cpu_numpy_baseline(n_cols, repeats=1):
    """The notebook beamformer, vectorised per channel, float64/complex128,
    all host cores through numpy's BLAS (torchrun exports OMP_NUM_THREADS=1: the limit is lifted here).
    Returns (samples/s, seconds, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy_beamformer as nb
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=cores)
    except Exception:
        limiter = None
    rng = np.random.default_rng(0xD5A)
    pos = np.zeros((N_ANT, 3)); pos[:, 0] = np.linspace(-250, 250, N_ANT)
    theta = np.linspace(-3.5 * np.pi / 180, 3.5 * np.pi / 180, N_BEAMS)
    A = nb.fourier_matrix(pos, theta, np.zeros(N_BEAMS))                                   # [chan][beam][ant]
    S = (rng.integers(-8, 8, size=(N_CHAN, N_ANT, n_cols)) + 1j * rng.integers(-8, 8, size=(N_CHAN, N_ANT, n_cols)))
    G = N_POL * N_AVG
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = np.empty((n_cols // G, N_CHAN, N_BEAMS))
        for k in range(N_CHAN):
            v = A[k] @ S[k]                                                                 # [beam][col]
            p = v.real ** 2 + v.imag ** 2
            out[:, k, :] = p.reshape(N_BEAMS, n_cols // G, G).sum(axis=2).T
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if limiter is not None:
        limiter.restore_original_limits()
    samples = N_BEAMS * N_CHAN * (n_cols // N_POL)
    return samples / best, best, cores


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    os.sched_setaffinity(0, range(os.cpu_count()))   # torchrun / numactl may have narrowed it: the CPU arm may use every core
    n_cols = 4096                        # 2048 input times x 2 pol per step: ~2 s of CPU work per step on 8 cores
    times = []
    cores = 1
    for i in range(args.warmup + args.steps):
        v, dt, cores = cpu_numpy_baseline(n_cols)
        if i >= args.warmup:
            times.append(dt)
    samples = N_BEAMS * N_CHAN * (n_cols // N_POL)
    value = samples * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d columns (%d input times x 2 pol) x 256 channels per step, numpy complex128 matmul per channel "
                                   "(port of sandbox/Beamformer Theory.ipynb; the reference has no compiled CPU path); one host, all of its "
                                   "cores, whatever --gpus says" % (n_cols, n_cols // 2)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU path
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx, pw = [], set(), None, []
        for r in rows:
            try:
                clk = float(r[1]); mx = float(r[2]); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hi = [c for c in sm if c >= 0.5 * max(sm)]          # samples taken under load
            out = {"sm_mhz": statistics.median(hi), "sm_min_mhz": min(hi), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                   "power_w_max": max(pw) if pw else None, "power_w_median": statistics.median(pw) if pw else None}
        return out


def int8_issue_peak():
    """Bare tcgen05.mma kind::i8 issue rate on this GPU (tools/ubench, ~1 s): the tensor-pipe ceiling no kernel with operands
    in shared memory can exceed.  MEASURED_PEAKS.json has no int8 entry; 2 x its bf16 figure is reported beside it."""
    exe = os.path.join(ROOT, "tools", "ubench")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    except (OSError, subprocess.TimeoutExpired):
        return None
    m = re.search(r"i8_mma_peak N=256:\s*[0-9.]+ ms\s+([0-9.]+) TOPS", r.stdout)
    return float(m.group(1)) if m else None


def reference_gpu_baseline(tmpdir):
    """The reference's own GPU sequence (expand_input -> cublasGemmStridedBatchedEx -> detect_sum) rebuilt for sm_100
    (oracle/_ref/ref_kernels_obs, built from /root/reference by oracle/Makefile), timed by its own CUDA events on this GPU.
    A reported baseline, not the target.  Returns None if the binary did not travel."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_kernels_obs")
    if not os.path.exists(exe):
        return None
    rng = np.random.default_rng(7)
    w = rng.integers(-127, 128, size=(N_CHAN, N_ANT, N_BEAMS, 2), dtype=np.int8)
    x = rng.integers(0, 256, size=(2, N_CHAN, 256, N_ANT), dtype=np.uint8)
    wf, xf = os.path.join(tmpdir, "w.i8"), os.path.join(tmpdir, "x.u8")
    w.tofile(wf); x.tofile(xf)
    try:
        r = subprocess.run([exe, wf, xf, "2", os.path.join(tmpdir, "ref"), "40"], capture_output=True, text=True, timeout=300)
    except (OSError, subprocess.TimeoutExpired):
        return None
    ms = {}
    for line in r.stdout.splitlines():
        parts = line.split()
        if len(parts) == 2 and parts[0].startswith("ref_ms_per_chunk"):
            ms[parts[0]] = float(parts[1])
    if not ms or "gemm_status 0" not in r.stdout:
        return None
    best = min(ms.values())
    samples_per_chunk = N_BEAMS * N_CHAN * 8 * N_AVG
    return {"samples_per_s": samples_per_chunk / (best * 1e-3), "ms_per_chunk_1_stream": ms.get("ref_ms_per_chunk"),
            "ms_per_chunk_8_streams": ms.get("ref_ms_per_chunk_8streams"),
            "what": "reference kernels + cuBLAS 12.9 C8I GEMM rebuilt with -arch=sm_100, device-resident, per 4 MiB chunk"}


def debug_shape_workload(bf, torch, device, weights):
    """BASELINE.json configs[1]-style shape (DEBUG build: N_AVERAGING = 1, 16 columns per chunk, 2 MiB of float power out
    per 256 KiB in): HBM-bound.  4096 chunks (= the 4096 grid sources) per launch, device-resident."""
    B = bf.Beamformer(debug=True, device=device, exact_power=0)
    B.set_weights(weights)
    n_chunks = 4096
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        x = torch.randint(0, 256, (n_chunks * B.chunk_bytes,), dtype=torch.uint8, device="cuda")
        out = torch.empty(n_chunks * B.chunk_out_floats, dtype=torch.float32, device="cuda")
        for _ in range(2):
            B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        reps = 5
        for _ in range(reps):
            B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n_chunks * (B.chunk_bytes + B.chunk_out_floats * 4)
    samples = N_BEAMS * N_CHAN * n_chunks * 8
    launches = B.kernel_launches
    B.close()
    del x, out
    return {"workload": "DEBUG shape (N_AVERAGING=1): 4096 chunks of 262144 B, 256 ch x 256 beams (configs[1] shape)", "ms_per_launch": ms,
            "samples_per_s": samples / (ms * 1e-3), "hbm_gb_s_algorithmic": byts / ms * 1e-6, "bound": "hbm"}, launches


def bind_near_gpu(torch, device):
    """Pin this rank's host threads to the CPUs of the GPU's NUMA node (sysfs local_cpulist) before any pinned
    buffer is allocated, so the staging memory is first-touched next to the GPU's PCIe root port -- the reference
    binds its process the same way (`-c core`, beamformer.cu:111-116).  Returns a short description."""
    try:
        pr = torch.cuda.get_device_properties(device)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unbound (no local cpulist)"
        os.sched_setaffinity(0, cpus)
        return "%d cpus local to %04x:%02x:%02x" % (len(cpus), pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except (OSError, AttributeError, ValueError) as e:
        return "unbound (%s)" % type(e).__name__


class Ranks:
    """barrier / reductions over the ranks of this run (no-ops for one GPU)"""

    def __init__(self, torch, world):
        self.torch, self.world = torch, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def min(self, x):
        return self._red(x, self.dist.ReduceOp.MIN) if self.dist else x

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def gather(self, x, rank):
        """x of every rank, in rank order (one all_reduce of a one-hot vector: the same primitive as max / sum)"""
        if not self.dist:
            return [float(x)]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device="cuda")
        t[rank] = float(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def copy_ceiling(torch, ranks, h_in, d_in, d_out, h_out, reps):
    """Bare pinned cudaMemcpyAsync of one step's bytes, H2D and D2H on two streams at once, on every rank at the same
    time: what the host side of this box can deliver to N GPUs with nothing else going on.  Returns GB/s per GPU."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)

    once()
    ranks.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = ranks.max(time.perf_counter() - t0)
    # H2D alone
    ranks.barrier()
    t1 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    dt_h2d = ranks.max(time.perf_counter() - t1)
    return {"h2d_gb_s_per_gpu_with_d2h": h_in.numel() * reps / dt * 1e-9, "d2h_gb_s_per_gpu_with_h2d": h_out.numel() * 4 * reps / dt * 1e-9,
            "h2d_gb_s_per_gpu_alone": h_in.numel() * reps / dt_h2d * 1e-9, "step_ms": 1e3 * dt / reps,
            "what": "pinned cudaMemcpyAsync of 1 GiB H2D || 512 MiB D2H per step on two streams, all ranks at once, host clock, max over ranks"}


def scaled_sharded_leg(bf, torch, ranks, rank, world, local_rank, peak_i8, peak_2bf16):
    """BASELINE.json configs[3]: 256 antennas x 2048 beams x 2048 channels, the channels split evenly over the ranks
    (strong scaling: total work fixed), 8 chunks (2048 columns = 1024 input times) per launch, device-resident,
    streamed-weights kernel.  Weights: a seeded random 128-channel slab repeated over the band."""
    n_ant, n_beams, n_chan_total, n_chunks = 256, 2048, 2048, 8
    c_local = n_chan_total // world
    B = bf.Beamformer(debug=False, device=local_rank, n_antennas=n_ant, n_beams=n_beams, n_channels=c_local, exact_power=0,
                      n_ring_slots=1, n_chunks_per_block=n_chunks)
    slab = np.random.default_rng(3 + rank).integers(-127, 128, size=(min(c_local, 128), n_ant, n_beams, 2), dtype=np.int8)
    B.set_weights(np.tile(slab, (c_local // slab.shape[0], 1, 1, 1)))
    del slab
    st = torch.cuda.Stream()
    gen = torch.Generator(device="cuda"); gen.manual_seed(77 + rank)
    with torch.cuda.stream(st):
        x = torch.randint(0, 256, (n_chunks * B.chunk_bytes,), dtype=torch.uint8, device="cuda", generator=gen)
        out = torch.empty(n_chunks * B.chunk_out_floats, dtype=torch.float32, device="cuda")
        for _ in range(2):
            B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        first = float(out[::4099].double().sum().item())
        ranks.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(st)
        for _ in range(reps):
            B.beamform_device(x.data_ptr(), n_chunks, out.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
    ms = ranks.max(e0.elapsed_time(e1) / reps)
    again = float(out[::4099].double().sum().item())
    ok = ranks.min(1.0 if (again == first and np.isfinite(first) and first > 0) else 0.0) == 1.0
    cols = n_chunks * B.cols_per_chunk
    samples = n_beams * n_chan_total * (cols // N_POL)
    ops = 8.0 * n_ant * n_beams * n_chan_total * cols
    launches = B.kernel_launches
    in_b, out_b = n_chunks * B.chunk_bytes, n_chunks * B.chunk_out_floats * 4
    B.close()
    del x, out
    tops = ops / ms * 1e-9
    return {"workload": "scaled array (configs[3]): 256 ant x 2048 beams x 2048 channels, %d channels per GPU, 8 chunks (1024 input times)" % c_local,
            "scaling": "strong", "channels_per_gpu": c_local, "ms_per_launch": ms, "samples_per_s": samples / (ms * 1e-3),
            "int8_tops_total": tops, "int8_tops_per_gpu": tops / world,
            "roofline": {"bound": "tensor", "achieved_per_gpu": tops / world, "peak": peak_i8, "frac": (tops / world / peak_i8) if peak_i8 else None,
                         "frac_of_2x_bf16_measured": tops / world / peak_2bf16, "unit": "TFLOP/s", "ops_per_sample": 8 * n_ant * N_POL},
            "bytes_in_per_gpu": in_b, "bytes_out_per_gpu": out_b,
            "check": "every rank: repeated launches bit-identical, finite, non-zero: %s; parity of this shape against the oracle: "
                     "tests/test_gpu_parity.py (scaled-array cases) and tools/mgpu_check.py (run by smoke() on > 1 GPU)" % ok}, launches


def dedisp_full_band_leg(bf, torch, ranks, rank, world, local_rank, hbm_peak):
    """BASELINE.json configs[4]: every rank holds detected power [T][256 channels][256 beams] of its sub-band; the
    frequency->beam exchange (ONE pull kernel over NVLink peer memory, csrc/exchange.cu; NCCL all-to-all if peer mapping
    is unavailable) gives [T][256*N channels][256/N beams]; then the 1627-trial DM sweep of sandbox/Dispersion
    Theory.ipynb (DM 0 ... 2000) on the local beams.  In-run check against plain torch (fp32): trial 0 = channel sum,
    the last trial = gather-and-sum."""
    from dsabeamformer_b200 import distributed as D
    T, C, NB = 2048, 256, 256
    B = bf.Beamformer(debug=True, n_channels=1, device=local_rank)
    gen = torch.Generator(device="cuda"); gen.manual_seed(900 + rank)
    freqs = np.concatenate([D.subband_frequencies_ghz(D.subband_of_rank(r)) for r in range(world)])
    dms = bf.dm_trials(2000.0, 2048)
    delay = bf.dm_delays(dms, freqs)
    n_t_out = T - int(delay.max())
    exch = {"kind": "none (one GPU holds every channel it has)", "ms": 0.0}
    hdl = None
    if world > 1:
        try:
            power, hdl = D.alloc_symmetric_power(T, C, NB)
            power.copy_(torch.rand((T, C, NB), dtype=torch.float32, device="cuda", generator=gen))
        except Exception as ex:                        # no peer mapping on this box: portable path
            hdl = None
            exch["fallback_reason"] = str(ex).split("\n")[0][:160]
            power = torch.rand((T, C, NB), dtype=torch.float32, device="cuda", generator=gen)
        torch.cuda.synchronize()
        times = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ranks.barrier()
            e0.record()
            full = D.freq_to_beam_pull(B, power, hdl) if hdl is not None else D.freq_to_beam_all_to_all(power)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = ranks.max(min(times[1:]))
        recv = T * C * (NB // world) * 4 * (world - 1)
        exch = dict(exch, kind=("pull kernel over NVLink peer memory (csrc/exchange.cu)" if hdl is not None else "NCCL all_to_all_single + 2 permute copies"),
                    ms=ms, bytes_received_per_gpu=recv, gb_s_received_per_gpu=recv / ms * 1e-6,
                    roofline={"bound": "nvlink", "achieved": recv / ms * 1e-6, "peak": 770.0, "unit": "GB/s", "frac": recv / ms * 1e-6 / 770.0,
                              "peak_source": "measured peer copy per direction per GPU, B200_PROFILING.md"})
        # every element must come from the right peer, channel and beam: compare with the NCCL transpose
        same = bool(torch.equal(full, D.freq_to_beam_all_to_all(power))) if hdl is not None else True
        exch["equals_nccl_all_to_all"] = ranks.min(1.0 if same else 0.0) == 1.0
    else:
        full = torch.rand((T, C, NB), dtype=torch.float32, device="cuda", generator=gen)
    Cf, Bl = C * world, NB // world
    out = torch.empty((len(dms), n_t_out, Bl), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    times = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ranks.barrier()
        with torch.cuda.stream(st):
            e0.record(st)
            B.dedisperse_device(full.data_ptr(), T, Cf, Bl, delay, n_t_out, out.data_ptr(), st.cuda_stream)
            e1.record(st)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = ranks.max(min(times[1:]))
    # in-run check with plain torch: DM 0 and the largest DM
    ref0 = full[:n_t_out].sum(dim=1)
    t_idx = torch.arange(n_t_out, device="cuda")[:, None] + torch.from_numpy(delay[-1].astype(np.int64)).cuda()[None, :]
    refl = full[t_idx, torch.arange(Cf, device="cuda")[None, :], :].sum(dim=1)
    worst = max(float(((out[0] - ref0).abs() / ref0).max().item()), float(((out[-1] - refl).abs() / refl).max().item()))
    ok = ranks.min(1.0 if worst <= 1e-5 else 0.0) == 1.0
    adds = float(len(dms)) * Cf * Bl * n_t_out
    byts = 4.0 * T * Cf * Bl + 4.0 * len(dms) * n_t_out * Bl
    launches = B.kernel_launches
    B.close()
    return {"workload": "DEBUG dedispersion (configs[4]): power [%d][256 ch][256 beams] per GPU -> exchange -> %d DM trials (0..2000) on [%d][%d ch][%d beams] per GPU"
                        % (T, len(dms), T, Cf, Bl),
            "exchange": exch,
            "dedisperse": {"n_dm": len(dms), "max_delay_samples": int(delay.max()), "n_t_out": n_t_out, "ms": ms, "adds_per_s_per_gpu": adds / ms * 1e3,
                           "adds_per_s_total": world * adds / ms * 1e3,
                           "roofline": {"bound": "hbm", "achieved": byts / ms * 1e-6, "peak": hbm_peak, "unit": "GB/s", "frac": byts / ms * 1e-6 / hbm_peak,
                                        "algorithmic_bytes": byts,
                                        "note": "brute force re-reads every (time, beam) window once per 8 trials from L2/shared memory; the binding "
                                                "unit at 1627 trials is the FADD/LDS issue rate (adds_per_s), not HBM"}},
            "check": "trial 0 and trial %d against torch (channel sum / gather-and-sum, fp32): max rel %.2e, ok on every rank: %s" % (len(dms) - 1, worst, ok)}, launches


def run_ours(args):
    import torch
    import dsabeamformer_b200 as bf
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py: --gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d bench.py ...)" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    host_binding = bind_near_gpu(torch, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ranks = Ranks(torch, world)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    bf16, bf16_sus = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    peak_2bf16 = 2.0 * bf16 if bf16 else 2.0 * 1590.0
    peak_2bf16_sus = 2.0 * bf16_sus if bf16_sus else 2.0 * 1400.0
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_i8 = int8_issue_peak() if rank == 0 else None          # ~1 s of tensor-pipe load, before anything is timed
    ranks.barrier()

    # --- set-up: one handle per GPU == one reference `beam -g rank` process
    B = bf.Beamformer(debug=False, device=local_rank, exact_power=0)
    pos = np.zeros((N_ANT, 3), dtype=np.float32)
    pos[:, 0] = np.array([i * 500.0 / (N_ANT - 1) - 250.0 for i in range(N_ANT)], dtype=np.float32)   # beamformer.cu:135-140
    dirs = np.zeros((N_BEAMS, 2), dtype=np.float32)
    PI = 3.14159265358979
    dirs[:, 0] = np.array([i * (7.0 * PI / 180.0) / (N_BEAMS - 1) - 3.5 * PI / 180.0 for i in range(N_BEAMS)], dtype=np.float32)
    B.set_weights(bf.compute_weights(pos, dirs, gpu=rank % 8))
    block_bytes, chunk_out = B.block_bytes, B.chunk_out_floats
    step_in_bytes = BLOCKS_PER_STEP * block_bytes
    step_out_bytes = BLOCKS_PER_STEP * CHUNKS_PER_BLOCK * chunk_out * 4
    samples_per_step = N_BEAMS * N_CHAN * BLOCKS_PER_STEP * CHUNKS_PER_BLOCK * 8 * N_AVG

    stream = torch.cuda.Stream()
    gen = torch.Generator(device="cuda"); gen.manual_seed(0xD5A + rank)
    with torch.cuda.stream(stream):
        d_in = torch.randint(0, 256, (step_in_bytes,), dtype=torch.uint8, device="cuda", generator=gen)   # junk bytes like dada_junkdb, fixed seed
        d_out = torch.empty(step_out_bytes // 4, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        for b in range(BLOCKS_PER_STEP):
            B.beamform_device(d_in.data_ptr() + b * block_bytes, CHUNKS_PER_BLOCK,
                              d_out.data_ptr() + b * CHUNKS_PER_BLOCK * chunk_out * 4, stream.cuda_stream)

    # --- value: device-resident, burst (K steps)
    for _ in range(args.warmup):
        step_device()
    ranks.barrier()
    sampler = ClockSampler(local_rank)
    launches0 = B.kernel_launches
    B.enable_kernel_timing(True)
    B.fused_kernel_ms(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    ranks.barrier()
    dev_ms = e0.elapsed_time(e1)
    kern_ms, kern_n = B.fused_kernel_ms(reset=True)
    B.enable_kernel_timing(False)
    launches_value = B.kernel_launches - launches0
    dev_ms_max = ranks.max(dev_ms)
    value = world * samples_per_step * args.steps / (dev_ms_max * 1e-3)
    # per rank: the step time and the kernel's own average launch time -- tells a slow GPU (both up) from a host that
    # fell behind with the launches (step up, kernel not); `value` uses the slowest rank
    per_rank = {"ms_per_step": [round(v / args.steps, 4) for v in ranks.gather(dev_ms, rank)],
                "fused_kernel_ms": [round(v, 5) for v in ranks.gather(kern_ms, rank)]}

    # --- e2e: host buffers through the streaming C ABI (H2D + kernel + D2H per block)
    h_in = torch.empty(step_in_bytes, dtype=torch.uint8).pin_memory()
    h_in.copy_(d_in.cpu())
    h_out = torch.empty(step_out_bytes // 4, dtype=torch.float32).pin_memory()
    n_slots = B.cfg.n_ring_slots

    def step_e2e():
        last = None
        for b in range(BLOCKS_PER_STEP):
            slot = b % n_slots
            B.enqueue_h2d(slot, h_in.data_ptr() + b * block_bytes, block_bytes)
            last = B.enqueue_block(slot, CHUNKS_PER_BLOCK, h_out.data_ptr() + b * CHUNKS_PER_BLOCK * chunk_out * 4)
        return last

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    B.sync()
    ranks.barrier()
    launches1 = B.kernel_launches
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    B.sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = B.kernel_launches - launches1
    ranks.barrier()
    clocks = sampler.stop()
    e2e_s_max = ranks.max(e2e_s)
    e2e_value = world * samples_per_step * args.steps / e2e_s_max
    # e2e result must equal the device-resident result bit for bit (same bytes, same kernel): the WHOLE step
    same = bool(torch.equal(h_out, d_out.cpu()))
    same_all = ranks.min(1.0 if same else 0.0) == 1.0
    ceiling = copy_ceiling(torch, ranks, h_in, d_in, d_out, h_out, max(3, min(args.steps, 10)))
    del h_in, h_out

    # --- sustained: the same device-resident loop for --sustain seconds without a gap
    sustained = None
    launches_sus = 0
    if args.sustain > 0:
        ranks.barrier()
        sus_sampler = ClockSampler(local_rank)
        launches2 = B.kernel_launches
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        n_steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.sustain:
            for _ in range(16):
                step_device()
            n_steps += 16
            if n_steps % 256 == 0:
                stream.synchronize()              # keep the host at most 256 steps ahead of the device
        s1.record(stream)
        ranks.barrier()
        sus_ms = ranks.max(s0.elapsed_time(s1))
        n_steps_min = ranks.min(float(n_steps))
        sus_clocks = sus_sampler.stop()
        launches_sus = B.kernel_launches - launches2
        sus_value = world * samples_per_step * n_steps_min / (sus_ms * 1e-3)      # conservative: fewest steps / longest time
        sus_tops_gpu = sus_value / world * OPS_PER_SAMPLE * 1e-12
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_steps, "value": sus_value, "unit": "samples/s", "ms_per_step": sus_ms / n_steps,
                     "int8_tops_per_gpu": sus_tops_gpu, "clocks": sus_clocks,
                     "roofline": {"bound": "tensor", "achieved": sus_tops_gpu, "peak": peak_2bf16_sus, "unit": "TFLOP/s", "frac": sus_tops_gpu / peak_2bf16_sus,
                                  "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained (seconds-long loop under the power cap)",
                                  "note": "whole loop (8 launches per step back to back), not the kernel alone"}}

    launches_total = int(ranks.sum(launches_value + launches_e2e + launches_sus))

    # --- the other BASELINE.json configs that need several GPUs to be what they are (also run at N = 1)
    extras = {}
    extra_launches = 0
    if not args.no_extras:
        peak_i8_all = ranks.max(peak_i8 or 0.0) or None
        extras["scaled_sharded"], n1 = scaled_sharded_leg(bf, torch, ranks, rank, world, local_rank, peak_i8_all, peak_2bf16)
        extras["dedisp_full_band"], n2 = dedisp_full_band_leg(bf, torch, ranks, rank, world, local_rank, hbm_peak)
        extra_launches = int(ranks.sum(n1 + n2))

    if rank == 0:
        peak = peak_i8 if peak_i8 else peak_2bf16
        peak_src = ("bare tcgen05.mma kind::i8 issue rate measured at the start of this run (tools/ubench, burst clocks); "
                    "MEASURED_PEAKS.json holds no int8 figure -- 2 x its bf16_tflops is frac_of_2x_bf16_measured" if peak_i8 else
                    ("2 x MEASURED_PEAKS.json bf16_tflops (tools/ubench did not run)" if bf16 else "2 x fallback 1.59 PFLOP/s of B200_PROFILING.md"))
        samples_per_launch = N_BEAMS * N_CHAN * CHUNKS_PER_BLOCK * 8 * N_AVG
        achieved = samples_per_launch * OPS_PER_SAMPLE / (kern_ms * 1e-3) * 1e-12 if kern_ms > 0 else 0.0
        algo_bytes = block_bytes + CHUNKS_PER_BLOCK * chunk_out * 4 + 2 * N_BEAMS * N_ANT * N_CHAN * 2
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], "static: %s (ncu --set full capture of this kernel, not measured by this run)" % tj.get("source", "profiles/")
        except (OSError, ValueError, KeyError):
            pass
        cpu_v, cpu_s, cores = cpu_numpy_baseline(4096, repeats=5)
        ref_gpu, dbg, dbg_launches = None, None, 0
        if not args.no_extras:
            with tempfile.TemporaryDirectory() as td:
                ref_gpu = reference_gpu_baseline(td)
            dbg, dbg_launches = debug_shape_workload(bf, torch, local_rank, bf.compute_weights(pos, dirs, gpu=rank % 8))
            dbg["frac_of_measured_hbm"] = dbg["hbm_gb_s_algorithmic"] / hbm_peak
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": CONFIG,
            "per_rank": per_rank,
            "value_sustained": sustained["value"] if sustained else None,
            "sustained": sustained,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": step_in_bytes, "d2h_bytes_per_step": step_out_bytes,
                    "ms_per_step": 1e3 * e2e_s_max / args.steps, "h2d_gb_s_per_gpu": step_in_bytes * args.steps / e2e_s_max * 1e-9,
                    "timing": "host clock between full device synchronisations (3 streams: H2D, compute, D2H)",
                    "matches_device_path": same_all, "host_binding": host_binding,
                    "copy_ceiling": ceiling,
                    "frac_of_copy_ceiling": ceiling["step_ms"] / (1e3 * e2e_s_max / args.steps)},
            "gpu_launches": launches_total + extra_launches + dbg_launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "frac_of_2x_bf16_measured": achieved / peak_2bf16, "peak_2x_bf16_measured": peak_2bf16,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "fused_beamform_kernel<32,0>", "kernel_ms": kern_ms, "launches_timed": kern_n,
                         "ops_per_launch": samples_per_launch * OPS_PER_SAMPLE, "algorithmic_bytes_per_launch": algo_bytes,
                         "peak_source": peak_src, "note": "int8 ops (TOPS); 8*N_ANT*N_POL = 1024 ops per beam x chan x input-time sample"},
            "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "4096 columns (2048 input times x 2 pol) x 256 channels, numpy complex128 matmul per channel, best of 5 passes of %.1f s" % cpu_s},
            "clocks": clocks,
            "reference_gpu_sm100": ref_gpu,
            "other_workloads": [w for w in (dbg,) if w],
            "scaled_sharded": extras.get("scaled_sharded"),
            "dedisp_full_band": extras.get("dedisp_full_band"),
            "x_realtime_per_gpu": value / world / REALTIME_SAMPLES_PER_S,
            "int8_tops": value * OPS_PER_SAMPLE * 1e-12,
        }
        print(json.dumps(line))
    B.close()
    if world > 1:
        ranks.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of continuous device-resident load for value_sustained (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip scaled_sharded, dedisp_full_band, the DEBUG shape and the reference GPU build")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
