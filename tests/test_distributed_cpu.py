"""CPU, world_size 2, gloo: the N>1 host logic -- channel/beam sharding and the frequency->beam
all-to-all that precedes full-band dedispersion (checked against the oracle on the gathered data)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsabeamformer_b200 import distributed as D
    import oracle as O
    T, C, B, n_dm, n_t_out = 40, 6, 8, 5, 24
    rng = np.random.default_rng(1234)                    # same stream on both ranks: the global truth
    full = rng.random((T, C * world, B), dtype=np.float32)
    c0, c1 = D.channel_range(rank, world, C * world)
    b0, b1 = D.beam_range(rank, world, B)
    mine = torch.from_numpy(np.ascontiguousarray(full[:, c0:c1, :]))
    got = D.freq_to_beam_all_to_all(mine).numpy()
    ok = np.array_equal(got, full[:, :, b0:b1])
    delay = rng.integers(0, 12, size=(n_dm, C * world)).astype(np.int32)
    ded_local = O.dedisperse(got, delay, n_t_out)        # what the GPU kernel computes on each rank
    ded_truth = O.dedisperse(full, delay, n_t_out)[:, :, b0:b1]
    ok = ok and np.allclose(ded_local, ded_truth, rtol=1e-6)
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([1.0 if ok else 0.0]))
    if rank == 0:
        ret.put(all(float(g) == 1.0 for g in gathered))
    dist.destroy_process_group()


def test_freq_to_beam_transpose_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_sharding_helpers():
    sys.path.insert(0, ROOT)
    from dsabeamformer_b200 import distributed as D
    assert [D.channel_range(r, 8, 2048) for r in (0, 7)] == [(0, 256), (1792, 2048)]
    assert D.beam_range(3, 4, 256) == (192, 256)
    assert D.subband_of_rank(5) == 5 and D.subband_of_rank(9) == 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    for g in (0, 3, 7):                                   # frequencies equal the oracle's weight frequencies (float quirks included)
        assert np.array_equal(D.subband_frequencies_ghz(g).astype(np.float32), O.freq_weights(g))


def test_bench_reference_arm_runs_under_two_ranks():
    """bench.py --impl reference: rank 0 prints one JSON line, other ranks exit 0 silently."""
    import json
    import subprocess
    env0 = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    env1 = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        env=env1, capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        env=env0, capture_output=True, text=True, timeout=300)
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["unit"] == "samples/s"
