#!/usr/bin/env python3
"""bench.py -- beamformed samples/s (beams x chans x input-time) of the fused
unpack+beamform+detect hot path, per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): the shape the reference's own number is quoted on
(README.md:11 -> BASELINE.md section 1) and BASELINE.json's configs[2]: 64 antennas
x 2 pol, 256 channels, 256 beams, N_AVERAGING = 16, PSRDADA-size blocks of
134 217 728 B (32 chunks), streamed.  One *step* = 8 such blocks (1 GiB of packed
voltages, > L2) through the hot path on every GPU; with N GPUs every rank owns one
256-channel sub-band (`-g rank`), no data-path collective (weak scaling).

  value : whole-job samples/s, inputs already resident in HBM, fused kernel only
          (dsabf_beamform_device), CUDA events on the launching stream, max over ranks
  e2e   : same metric through the C ABI with HOST buffers: pinned H2D of every block
          into the device ring, fused kernel, D2H of the float power into pinned memory
          (dsabf_enqueue_h2d / dsabf_enqueue_block), all inside the timed region
  roofline : dominant kernel = fused_beamform_kernel; achieved = algorithmic int8 ops
          (8*N_ANT*N_POL = 1024 per sample) / mean kernel time (CUDA events on its stream)
  cpu_baseline : the reference's numpy beamformer (oracle/numpy_beamformer.py, a port of
          sandbox/Beamformer Theory.ipynb) on this host's cores, bounded sample
--impl reference times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
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


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------- CPU path
def cpu_numpy_baseline(n_cols, repeats=1):
    """The notebook beamformer, vectorised per channel, float64/complex128,
    all host cores through numpy's BLAS.  Returns (samples/s, seconds, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy_beamformer as nb
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
    samples = N_BEAMS * N_CHAN * (n_cols // N_POL)
    return samples / best, best, os.cpu_count()


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    n_cols = 4096                        # 2048 input times x 2 pol per step: ~2 s of CPU work per step on 8 cores
    times = []
    for i in range(args.warmup + args.steps):
        v, dt, cores = cpu_numpy_baseline(n_cols)
        if i >= args.warmup:
            times.append(dt)
    samples = N_BEAMS * N_CHAN * (n_cols // N_POL)
    value = samples * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "beamformed samples/s (beams x chans x input-time)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_antennas": N_ANT, "n_beams": N_BEAMS, "n_channels": N_CHAN, "n_averaging": N_AVG},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d columns (%d input times x 2 pol) x 256 channels per step, numpy complex128 matmul per channel "
                                   "(port of sandbox/Beamformer Theory.ipynb; the reference has no compiled CPU path)" % (n_cols, n_cols // 2)},
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
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                   "power_w_max": max(pw) if pw else None}
        return out



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


def scaled_array_workload(bf, torch, device):
    """BASELINE.json configs[3], one GPU's share of the frequency-sharded scaled array: 256 antennas x 2048 beams x
    256 channels, one 32-chunk block (8192 columns, 512 MiB in / 512 MiB out), device-resident, random int8 weights
    (streamed-weights kernel, fused_beamform_big.cu): tensor-bound."""
    import numpy as np
    n_ant, n_beams, n_chan, n_chunks = 256, 2048, 256, 32
    B = bf.Beamformer(debug=False, device=device, n_antennas=n_ant, n_beams=n_beams, n_channels=n_chan, exact_power=0)
    B.set_weights(np.random.default_rng(3).integers(-127, 128, size=(n_chan, n_ant, n_beams, 2), dtype=np.int8))
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
    cols = n_chunks * B.cols_per_chunk
    samples = n_beams * n_chan * (cols // N_POL)
    launches = B.kernel_launches
    B.close()
    del x, out
    return {"workload": "scaled array (configs[3], one GPU's 256 of 2048 channels): 256 ant x 2048 beams x 256 ch, 32 chunks", "ms_per_launch": ms,
            "samples_per_s": samples / (ms * 1e-3), "int8_tops": 8.0 * n_ant * n_beams * n_chan * cols / ms * 1e-9, "bound": "tensor"}, launches


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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

    # --- value: device-resident
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = B.kernel_launches
    B.enable_kernel_timing(True)
    B.fused_kernel_ms(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    kern_ms, kern_n = B.fused_kernel_ms(reset=True)
    B.enable_kernel_timing(False)
    launches_value = B.kernel_launches - launches0
    dev_ms_max = max_over_ranks(dev_ms)
    value = world * samples_per_step * args.steps / (dev_ms_max * 1e-3)

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
    barrier()
    launches1 = B.kernel_launches
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    B.sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = B.kernel_launches - launches1
    barrier()
    clocks = sampler.stop()
    e2e_s_max = max_over_ranks(e2e_s)
    e2e_value = world * samples_per_step * args.steps / e2e_s_max
    # e2e result must equal the device-resident result bit for bit (same bytes, same kernel)
    same = bool(torch.equal(h_out[: chunk_out * 4], d_out[: chunk_out * 4].cpu()))

    launches_total = int(sum_over_ranks(launches_value + launches_e2e))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        bf16 = peaks.get("bf16_tflops")
        peak = 2.0 * bf16 if bf16 else 2.0 * 1590.0
        peak_src = ("2 x MEASURED_PEAKS.json bf16_tflops (tcgen05 kind::i8 = 2x kind::f16 rate), burst figure: kernel timed alone"
                    if bf16 else "2 x fallback 1.59 PFLOP/s of B200_PROFILING.md (MEASURED_PEAKS.json absent)")
        samples_per_launch = N_BEAMS * N_CHAN * CHUNKS_PER_BLOCK * 8 * N_AVG
        achieved = samples_per_launch * OPS_PER_SAMPLE / (kern_ms * 1e-3) * 1e-12 if kern_ms > 0 else 0.0
        algo_bytes = block_bytes + CHUNKS_PER_BLOCK * chunk_out * 4 + 2 * N_BEAMS * N_ANT * N_CHAN * 2
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")))["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        cpu_v, cpu_s, cores = cpu_numpy_baseline(4096, repeats=5)
        with tempfile.TemporaryDirectory() as td:
            ref_gpu = reference_gpu_baseline(td)
        dbg, dbg_launches = debug_shape_workload(bf, torch, local_rank, bf.compute_weights(pos, dirs, gpu=rank % 8))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        dbg["frac_of_measured_hbm"] = dbg["hbm_gb_s_algorithmic"] / hbm_peak
        big, big_launches = scaled_array_workload(bf, torch, local_rank)
        big["frac_of_peak"] = big["int8_tops"] / peak
        line = {
            "metric": "beamformed samples/s (beams x chans x input-time)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_antennas": N_ANT, "n_pol": N_POL, "n_channels_per_gpu": N_CHAN, "n_beams": N_BEAMS,
                       "n_averaging": N_AVG, "blocks_per_step": BLOCKS_PER_STEP, "bytes_in_per_step_per_gpu": step_in_bytes,
                       "l2_policy": "inputs larger than L2 (1 GiB of packed voltages per step per GPU, 512 MiB of power out)",
                       "parallelism": "channel sub-band per GPU, no collective", "power_mode": "fp32 (bit-exact below 2^24)"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": step_in_bytes, "d2h_bytes_per_step": step_out_bytes,
                    "ms_per_step": 1e3 * e2e_s_max / args.steps, "h2d_gb_s_per_gpu": step_in_bytes * args.steps / e2e_s_max * 1e-9,
                    "timing": "host clock between full device synchronisations (3 streams: H2D, compute, D2H)",
                    "matches_device_path": same, "host_binding": host_binding},
            "gpu_launches": launches_total,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "fused_beamform_kernel<32,0>", "kernel_ms": kern_ms, "launches_timed": kern_n,
                         "ops_per_launch": samples_per_launch * OPS_PER_SAMPLE, "algorithmic_bytes_per_launch": algo_bytes,
                         "peak_source": peak_src, "note": "int8 ops (TOPS); 8*N_ANT*N_POL = 1024 ops per beam x chan x input-time sample"},
            "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "4096 columns (2048 input times x 2 pol) x 256 channels, numpy complex128 matmul per channel, best of 5 passes of %.1f s" % cpu_s},
            "clocks": clocks,
            "reference_gpu_sm100": ref_gpu,
            "other_workloads": [dbg, big],
            "x_realtime_per_gpu": value / world / REALTIME_SAMPLES_PER_S,
            "int8_tops": value * OPS_PER_SAMPLE * 1e-12,
        }
        print(json.dumps(line))
    B.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
