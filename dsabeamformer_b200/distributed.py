"""Multi-GPU plumbing: channel sharding and the one exchange the path has.

Unpack + beamform + detect shard by frequency with no communication: the reference already runs
8 independent `beam -g <0..7>` processes of 256 of the 2048 channels (/root/reference/src/
beamformer.hh:63-64; README.md:168).  Only full-band dedispersion needs data from every channel,
so the detected power is transposed from "all beams of my channels" to "all channels of my beams"
with ONE all-to-all (NCCL over NVLink on GPUs; gloo in the CPU tests), then dedispersed locally.
torch.distributed is the plumbing; the kernels are the library's.
"""
import numpy as np
import torch
import torch.distributed as dist

TOT_CHANNELS, N_SUBBANDS = 2048, 8


def subband_of_rank(rank):
    """`-g` of the reference: rank r beamforms sub-band r (integer offset r*2048/7 channels, beamformer.cu:233)."""
    return rank % N_SUBBANDS


def channel_range(rank, world, n_channels_total):
    """Contiguous channel shard of `rank` (equal shards; n_channels_total % world == 0)."""
    assert n_channels_total % world == 0
    per = n_channels_total // world
    return rank * per, (rank + 1) * per


def beam_range(rank, world, n_beams):
    assert n_beams % world == 0
    per = n_beams // world
    return rank * per, (rank + 1) * per


def subband_frequencies_ghz(gpu, n_chan=256):
    """Channel centre frequencies of sub-band `gpu` as the steering weights see them (beamformer.cu:233)."""
    bw = np.float32((1.53 - 1.28) / TOT_CHANNELS)
    idx = gpu * TOT_CHANNELS // (N_SUBBANDS - 1) + np.arange(n_chan)
    return (1.53 - (idx.astype(np.float32) * bw).astype(np.float64)).astype(np.float32).astype(np.float64)


def freq_to_beam_all_to_all(power_local, group=None):
    """power_local: [T][ch_local][B] (all beams, my channels) on every rank  ->  [T][ch_local*world][B/world]
    (all channels in rank order, my beams).  One all_to_all_single; message to each peer is
    T*ch_local*(B/world) floats."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    T, C, B = power_local.shape
    assert B % world == 0
    Bs = B // world
    if world == 1:
        return power_local
    # [T][C][world][Bs] -> [world][T][C][Bs]: chunk h goes to rank h
    send = power_local.reshape(T, C, world, Bs).permute(2, 0, 1, 3).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[g] = rank g's channels for my beams -> [T][g*C + c][Bs]
    return recv.permute(1, 0, 2, 3).reshape(T, world * C, Bs).contiguous()


def alloc_symmetric_power(n_t, c_local, n_beams, group=None):
    """Power buffer [n_t][c_local][n_beams] that every rank of `group` can read directly over NVLink
    (torch symmetric memory: plumbing only).  Returns (tensor, handle); handle.buffer_ptrs are the peers'
    device pointers as mapped on this GPU.  Collective: every rank must call it with the same shape."""
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((n_t, c_local, n_beams), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    return t, hdl


def freq_to_beam_pull(beamformer, power_symm, hdl):
    """Same result as freq_to_beam_all_to_all, but as ONE kernel of the library that pulls this rank's beams of
    every peer's channels straight out of the peers' memory (csrc/exchange.cu) -- no permute copies, no NCCL.
    power_symm / hdl come from alloc_symmetric_power.  Fenced by device-side barriers on the symmetric
    allocation's signal pads: peers have finished writing before the pull, and are released after it."""
    T, C, B = power_symm.shape
    world, rank = hdl.world_size, hdl.rank
    full = torch.empty((T, C * world, B // world), dtype=torch.float32, device=power_symm.device)
    cur = torch.cuda.current_stream(power_symm.device)
    side = _side_stream(power_symm.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        hdl.barrier(channel=0)
        beamformer.transpose_pull(list(hdl.buffer_ptrs), rank, T, C, B, full.data_ptr(), side.cuda_stream)
        hdl.barrier(channel=1)
    cur.wait_stream(side)
    full.record_stream(side)
    return full


def dedisperse_single_full_band_fused(beamformer, power_symm, hdl, delay_1d, n_t_out):
    """One trial (DM 0, or one delay curve: int32 [ch_local*world]) over the full band with the exchange FUSED into
    the sum: the library's streaming kernel reads this rank's beams of every peer's channels in place over NVLink
    (csrc/dedisperse.cu, dsabf_dedisperse_single_peers_device).  Returns [n_t_out][B/world]."""
    T, C, B = power_symm.shape
    world, rank = hdl.world_size, hdl.rank
    out = torch.empty((n_t_out, B // world), dtype=torch.float32, device=power_symm.device)
    cur = torch.cuda.current_stream(power_symm.device)
    side = _side_stream(power_symm.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        hdl.barrier(channel=0)
        beamformer.dedisperse_single_peers(list(hdl.buffer_ptrs), rank, T, C, B, delay_1d, n_t_out, out.data_ptr(), side.cuda_stream)
        hdl.barrier(channel=1)
    cur.wait_stream(side)
    out.record_stream(side)
    return out


def dedisperse_full_band(beamformer, power_local, delay_full, n_t_out, group=None, symm_handle=None):
    """Full-band incoherent dedispersion across ranks.  delay_full: int32 [n_dm][ch_local*world] in the
    global channel order (rank-major).  Returns this rank's beams: [n_dm][n_t_out][B/world] (CUDA tensor).
    With symm_handle (power_local allocated by alloc_symmetric_power) the exchange is the library's own
    pull kernel over NVLink peer memory; otherwise NCCL all_to_all_single (gloo on CPU)."""
    full = freq_to_beam_pull(beamformer, power_local, symm_handle) if symm_handle is not None else freq_to_beam_all_to_all(power_local, group)
    T, C, Bs = full.shape
    n_dm = delay_full.shape[0]
    out = torch.empty((n_dm, n_t_out, Bs), dtype=torch.float32, device=full.device)
    # The C ABI takes a raw stream handle and treats 0 as "use the handle's own compute stream", which is not
    # ordered after torch's legacy default stream: run the kernel on a real side stream fenced on both ends.
    cur = torch.cuda.current_stream(full.device)
    side = _side_stream(full.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        beamformer.dedisperse_device(full.data_ptr(), T, C, Bs, np.ascontiguousarray(delay_full, dtype=np.int32), n_t_out, out.data_ptr(),
                                     side.cuda_stream)
    cur.wait_stream(side)
    full.record_stream(side)
    out.record_stream(side)
    return out


_side_streams = {}


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return _side_streams[key]
