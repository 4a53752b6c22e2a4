"""GPU, N > 1: sub-band sharding and the NCCL frequency->beam all-to-all + dedispersion, under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_subbands_and_full_band_dedispersion():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("vs oracle") == 4 * world          # sub-band, full-band dedispersion, scaled-array shard, full trial list per rank
