#!/usr/bin/env python3
"""Time the dedispersion kernel on one GPU: [T][2048 ch][B] power -> [n_dm][T_out][B]; report achieved
GB/s on the algorithmic bytes (read once + write) and effective adds/s."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dsabeamformer_b200 as bf
import oracle as O

B = bf.Beamformer(debug=True, n_channels=1)
n_t, n_chan, n_beams = int(os.environ.get("DD_T", 1024)), 2048, int(os.environ.get("DD_B", 32))
n_dm = int(os.environ.get("DD_NDM", 256))
freq = 1.53 - (np.arange(n_chan) + 0.0) * (0.25 / 2048)
stride = int(os.environ.get("DD_STRIDE", 1))        # 1: neighbouring trials of the list (equal delays at most channels); > 1: every stride-th
dms = O.dm_trials(2000.0, 2048)[::stride][:n_dm]
delay = O.dm_delays(dms, freq)
n_t_out = n_t - int(delay.max())
assert n_t_out > 0, (n_t, delay.max())
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    power = torch.rand((n_t, n_chan, n_beams), dtype=torch.float32, device="cuda")
    out = torch.empty((n_dm, n_t_out, n_beams), dtype=torch.float32, device="cuda")
    for _ in range(2):
        B.dedisperse_device(power.data_ptr(), n_t, n_chan, n_beams, delay, n_t_out, out.data_ptr(), st.cuda_stream)
    reps, batches = 10, []
    for _ in range(5):                      # 5 batches of 10 launches: report the median batch (run-to-run spread is large)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            B.dedisperse_device(power.data_ptr(), n_t, n_chan, n_beams, delay, n_t_out, out.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        batches.append(e0.elapsed_time(e1) / reps)
ms = sorted(batches)[len(batches) // 2]
algo = 4.0 * n_t * n_chan * n_beams + 4.0 * n_dm * n_t_out * n_beams
adds = float(n_dm) * n_chan * n_beams * n_t_out
# spot check vs oracle on a slice of beams / trials
sub = O.dedisperse(power[:, :, :2].cpu().numpy(), delay[:3], n_t_out)
got = out[:3, :, :2].cpu().numpy()
print(json.dumps({"kernel": "dedisperse_kernel", "n_t": n_t, "n_chan": n_chan, "n_beams": n_beams, "n_dm": n_dm, "n_t_out": n_t_out,
                  "max_delay": int(delay.max()), "dm_stride": stride, "ms": ms, "ms_batches": [round(b, 4) for b in batches], "algorithmic_GB_s": algo / ms * 1e-6, "Gadds_s": adds / ms * 1e-6,
                  "max_rel_vs_oracle": float((np.abs(got - sub) / np.maximum(np.abs(sub), 1e-30)).max())}))
