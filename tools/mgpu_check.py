#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
  * every rank beamforms its own sub-band (`-g rank`) of the same sources -> matches the oracle for that sub-band
  * full-band dedispersion: NCCL all-to-all (frequency -> beam transpose) + local dedisperse kernel ->
    matches the oracle run on the gathered full-band power
Prints one line per check; exit code != 0 on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dsabeamformer_b200 as bf  # noqa: E402
from dsabeamformer_b200 import distributed as D  # noqa: E402
import make_configs  # noqa: E402
import oracle as O  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfgdir = os.path.join(ROOT, "config")
if rank == 0 and not os.path.exists(os.path.join(cfgdir, "grid_positions.txt")):
    make_configs.make_all(cfgdir)
dist.barrier()
pos = O.read_positions(os.path.join(cfgdir, "grid_positions.txt"))
dirs = O.read_directions(os.path.join(cfgdir, "grid_beam_directions.txt"), 256)
src = O.read_directions(os.path.join(cfgdir, "grid_source_directions_4096.txt"))
gpu = D.subband_of_rank(rank)
n_chunks = 6
B = bf.Beamformer(debug=False, device=local)
w = bf.compute_weights(pos, dirs, gpu)
B.set_weights(w)
data = O.generate(pos, src[[7, 900, 2048, 2100, 3000, 4095]], n_chunks=n_chunks, cols=B.cols_per_chunk, gpu=gpu)
power = B.beamform(data)                                           # [48][256][256]
ok = True
ref = O.detect_ref_float(O.beamform(data[2], O.weights(pos, dirs, gpu)), 32)
rel = np.abs(power[16:24] - ref) / np.maximum(ref, 1e-30)
ok &= bool(rel.max() <= 1e-6)
print("rank %d sub-band %d beamform vs oracle: max rel %.2e" % (rank, gpu, rel.max()), flush=True)

# full-band dedispersion over world*256 channels
T = power.shape[0]
d_power = torch.from_numpy(power).cuda()
freqs = np.concatenate([D.subband_frequencies_ghz(D.subband_of_rank(r)) for r in range(world)])
dms = O.dm_trials(100.0, 2048)[::3][:16]
delay = np.minimum(O.dm_delays(dms, freqs), 20).astype(np.int32)
n_t_out = T - 20
ded = D.dedisperse_full_band(B, d_power, delay, n_t_out).cpu().numpy()       # [n_dm][n_t_out][256/world]
gathered = [torch.empty_like(d_power) for _ in range(world)]
dist.all_gather(gathered, d_power)
full = torch.cat(gathered, dim=1).cpu().numpy()                              # [T][world*256][256]
b0, b1 = D.beam_range(rank, world, 256)
truth = O.dedisperse(np.ascontiguousarray(full[:, :, b0:b1]), delay, n_t_out)
rel = np.abs(ded - truth) / np.maximum(np.abs(truth), 1e-30)
# float32 sums of world*256 positive terms in two different orders (kernel: 4-channel chunks; oracle: sequential):
# typical difference sqrt(n)*2^-24, allowed n*1e-8
ok &= bool(rel.max() <= max(5e-6, 1e-8 * full.shape[1]))
try:        # same dedispersion with the exchange done by the library's NVLink pull kernel: bit-identical
    sp, sh = D.alloc_symmetric_power(*d_power.shape)
    sp.copy_(d_power)
    ded2 = D.dedisperse_full_band(B, sp, delay, n_t_out, symm_handle=sh).cpu().numpy()
    same = bool(np.array_equal(ded, ded2))
    ok &= same
    print("rank %d full-band dedispersion with the NVLink pull exchange identical to the NCCL one: %s" % (rank, same), flush=True)
except Exception as ex:
    print("rank %d symmetric memory unavailable (%s): NCCL exchange only" % (rank, str(ex).split("\n")[0][:120]), flush=True)
print("rank %d full-band dedispersion (all-to-all + kernel) vs oracle: max rel %.2e" % (rank, rel.max()), flush=True)

# BASELINE.json configs[3] (scaled array, frequency-sharded): 256 antennas x 2048 beams, this rank's channels
rng = np.random.default_rng(500 + rank)
Bs = bf.Beamformer(debug=False, device=local, n_antennas=256, n_beams=2048, n_channels=3)
ws = rng.integers(-127, 128, size=(3, 256, 2048, 2), dtype=np.int8)
Bs.set_weights(ws)
xs = rng.integers(0, 256, size=(2, 3, Bs.cols_per_chunk, 256), dtype=np.uint8)
ps = Bs.beamform(xs)
worst = 0.0
for c in range(2):
    ref = O.detect_ref_float(O.beamform(xs[c], ws), 32)
    worst = max(worst, float((np.abs(ps[c * 8:(c + 1) * 8] - ref) / np.maximum(ref, 1e-30)).max()))
ok &= bool(worst <= 1e-6)
print("rank %d scaled-array shard (256 ant x 2048 beams) vs oracle: max rel %.2e" % (rank, worst), flush=True)
Bs.close()

# BASELINE.json configs[4] as specified: the WHOLE trial list of sandbox/Dispersion Theory.ipynb (1627 trials, DM 0..2000,
# unclamped delays: several hundred samples across the band) on the exchanged full band, every 50th trial and the
# last one against the oracle
dms_all = bf.dm_trials(2000.0, 2048)
delay_all = bf.dm_delays(dms_all, freqs)
T2 = int(delay_all.max()) + 96
gen = torch.Generator(device="cuda"); gen.manual_seed(4242 + rank)
try:
    p2, h2 = D.alloc_symmetric_power(T2, 256, 256)
    p2.copy_(torch.rand((T2, 256, 256), dtype=torch.float32, device="cuda", generator=gen) * 10)
except Exception:
    p2, h2 = torch.rand((T2, 256, 256), dtype=torch.float32, device="cuda", generator=gen) * 10, None
torch.cuda.synchronize(); dist.barrier()
ded_all = D.dedisperse_full_band(B, p2, delay_all, 96, symm_handle=h2).cpu().numpy()          # [1627][96][256/world]
g2 = [torch.empty_like(p2) for _ in range(world)]
dist.all_gather(g2, p2)
full2 = torch.cat(g2, dim=1)[:, :, b0:b1].contiguous().cpu().numpy()
del g2
pick = np.unique(np.concatenate([np.arange(0, len(dms_all), 50), [len(dms_all) - 1]]))
truth2 = O.dedisperse(full2, np.ascontiguousarray(delay_all[pick]), 96)
rel2 = np.abs(ded_all[pick] - truth2) / np.maximum(np.abs(truth2), 1e-30)
ok &= bool(rel2.max() <= max(5e-6, 1e-8 * full2.shape[1]))
print("rank %d full trial list (%d trials, max delay %d samples, %s exchange) vs oracle on %d sampled trials: max rel %.2e"
      % (rank, len(dms_all), int(delay_all.max()), "NVLink pull" if h2 is not None else "NCCL", len(pick), rel2.max()), flush=True)
del full2, ded_all, p2

# timing of the one exchange the path has, at the shard shape of configs[4]: [1024 times][256 ch][256 beams] per rank
# -> all-to-all -> [1024][256*world][256/world], then 256 DM trials on the local beams (CUDA events, max over ranks)
if world > 1:
    Tt = 1024
    pw = torch.rand((Tt, 256, 256), dtype=torch.float32, device="cuda")
    dms_t = O.dm_trials(2000.0, 2048)[:256]
    delay_t = O.dm_delays(dms_t, np.concatenate([D.subband_frequencies_ghz(D.subband_of_rank(r)) for r in range(world)])).astype(np.int32)
    nto = Tt - int(delay_t.max())
    times = []
    ts = torch.cuda.Stream()
    out_t = torch.empty((256, nto, 256 // world), dtype=torch.float32, device="cuda")
    for rep in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        dist.barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(ts):
            e[0].record()
            full_t = D.freq_to_beam_all_to_all(pw)
            e[1].record()
            B.dedisperse_device(full_t.data_ptr(), Tt, 256 * world, 256 // world, delay_t, nto, out_t.data_ptr(), ts.cuda_stream)
            e[2].record()
        torch.cuda.synchronize()
        times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    # the same transpose as one kernel of ours over NVLink peer memory (csrc/exchange.cu)
    pull_ms, pull_ok = float("nan"), True
    try:
        pws, hdl = D.alloc_symmetric_power(Tt, 256, 256)
        pws.copy_(pw)
        torch.cuda.synchronize(); dist.barrier()
        ref_full = D.freq_to_beam_all_to_all(pw)
        pt = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            got_full = D.freq_to_beam_pull(B, pws, hdl)
            e1.record(); torch.cuda.synchronize()
            pt.append(e0.elapsed_time(e1))
        pull_ok = bool(torch.equal(got_full, ref_full))
        ok &= pull_ok
        pull_ms = min(pt[1:])
        print("rank %d NVLink pull transpose == NCCL all-to-all: %s (%.3f ms)" % (rank, pull_ok, pull_ms), flush=True)
        # full-band DM 0 (the reference's gemv over ALL sub-bands): exchange fused into the streaming sum vs transpose-then-sum
        zero = np.zeros((1, 256 * world), dtype=np.int32)
        two_step = D.dedisperse_full_band(B, pws, zero, Tt, symm_handle=hdl)[0]
        ft = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            fused = D.dedisperse_single_full_band_fused(B, pws, hdl, zero[0], Tt)
            e1.record(); torch.cuda.synchronize()
            ft.append(e0.elapsed_time(e1))
        same0 = bool(torch.equal(fused, two_step))
        ok &= same0
        print("rank %d full-band DM 0 fused over NVLink == transpose + sum: %s (%.3f ms)" % (rank, same0, min(ft[1:])), flush=True)
    except Exception as ex:                                     # symmetric memory unavailable on this box: NCCL path only
        print("rank %d NVLink pull transpose unavailable: %s" % (rank, str(ex).split("\n")[0][:160]), flush=True)
    tt = torch.tensor(times[1:], dtype=torch.float64, device="cuda").min(dim=0).values
    pm = torch.tensor([pull_ms if pull_ms == pull_ms else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        sent = Tt * 256 * 256 * 4 * (world - 1) / world
        adds = 256.0 * 256 * world * (256 // world) * nto
        if pm.item() > 0:
            print("rank 0 timing: transpose as one pull kernel over NVLink peer memory %.3f ms = %.1f GB/s received per GPU" % (pm.item(), sent / pm.item() * 1e-6), flush=True)
        print("rank 0 timing: transpose (2 permute copies + NCCL all_to_all_single) %.3f ms = %.1f GB/s sent per GPU; dedisperse 256 trials x %d ch x %d beams %.3f ms = %.2e adds/s per GPU, %.2e aggregate"
              % (tt[0].item(), sent / tt[0].item() * 1e-6, 256 * world, 256 // world, tt[1].item(), adds / tt[1].item() * 1e3, world * adds / tt[1].item() * 1e3), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
B.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
