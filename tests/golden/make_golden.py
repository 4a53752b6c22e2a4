#!/usr/bin/env python3
"""Generate tests/golden/ref_host_golden.json and ref_host_small.npz.

Runs the REFERENCE's own host expressions (oracle/_ref/ref_host, compiled by
oracle/Makefile from /root/reference/src/{beamformer.hh,beamformer.cuh,
test_data_generator.hh} and the weight loop of src/beamformer.cu:232-241) on the
reference's own config files, and records
  * sha256 of the full steering-weight table and of every generated 256 MiB
    test-data batch, per configuration,
  * a small slice of each (weights of channels 0,255; chunk/source 0,100,1023
    channels 0,255) verbatim,
so that tests can pin the oracle (and the product's host code) to the reference
without /root/reference being present.  Only runs where /root/reference exists.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "ref_host")

CASES = [
    # name, positions, directions, sources, gpu, batches
    ("linear_g0", "linear_positions.txt", "linear_directions.txt", "linear_source_directions_1024.txt", 0, 1),
    ("grid4096_g0", "grid_positions.txt", "grid_beam_directions.txt", "grid_source_directions_4096.txt", 0, 4),
    ("grid3721_g3", "grid_positions.txt", "grid_beam_directions.txt", "grid_source_directions_3721.txt", 3, 4),
    ("defaults_bogus_g5", None, None, None, 5, 1),
]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    golden, small = {}, {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, p, d, s, gpu, nb in CASES:
            cfg = lambda f: os.path.join(REF, "config", f) if f else "-"
            prefix = os.path.join(tmp, name)
            subprocess.check_call([REF_HOST, cfg(p), cfg(d), cfg(s), str(gpu), prefix, str(nb)], stdout=subprocess.DEVNULL)
            w = np.fromfile(prefix + ".weights.i8", dtype=np.int8).reshape(256, 64, 256, 2)
            entry = {"positions": p, "directions": d, "sources": s, "gpu": gpu,
                     "weights_sha256": hashlib.sha256(w.tobytes()).hexdigest(), "data_sha256": []}
            small[name + "/w_ch0"] = w[0]
            small[name + "/w_ch255"] = w[255]
            for b in range(nb):
                data = np.fromfile(prefix + ".data.b%d.u8" % b, dtype=np.uint8).reshape(1024, 256, 16, 64)
                entry["data_sha256"].append(hashlib.sha256(data.tobytes()).hexdigest())
                # column 0 only (all 16 columns of a chunk are identical: test_data_generator.hh:76-84)
                small[name + "/data_b%d" % b] = data[[0, 100, 1023]][:, [0, 255], 0, :]
                os.remove(prefix + ".data.b%d.u8" % b)
            golden[name] = entry
    # known answers the reference itself contains
    golden["nibble_kat"] = {"0x25": [2, 5], "0xA8": [-6, -8], "0xD7": [-3, 7], "0x70": [7, 0]}   # sandbox/bitshift.cpp:5-28, kernelTest.cu:127-148, test_data_generator.hh:8
    golden["notebook_sum_A_linear"] = 13295.149606299225    # sandbox/2D Beamformer.ipynb cell 6 stored output
    golden["dm_trials_to_2000"] = 1627                       # sandbox/Dispersion Theory.ipynb cell 2 stored output
    with open(os.path.join(HERE, "ref_host_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "ref_host_small.npz"), **small)
    print("wrote golden for", list(golden))


if __name__ == "__main__":
    main()
