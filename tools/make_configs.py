#!/usr/bin/env python3
"""Regenerate the array / beam / source configuration files under config/.

The reference ships its inputs as text files (format: SURVEY.md section 5.6;
parsers at /root/reference/src/beamformer.hh:250-284 and
src/test_data_generator.hh:43-60).  They are all formula-generated grids, so we
regenerate them here instead of copying them; tests/test_oracle_golden.py
(test_generated_configs_parse_like_reference) checks, where
the reference is mounted, that every value parses to the *same float32* as the
reference's file of the same name.

    positions : first token = count, then "x y z" per antenna, metres
    directions: first token = count, then "theta phi" per beam/source, radians

random_positions.txt is the one exception: the reference's file of that name is an
unseeded random draw (sandbox notebook), cannot be regenerated and is not used by any
named configuration, so ours is a fixed-seed draw of the same kind (64 antennas uniform in
+-250 m) under the same name -- NOT the same numbers.  The reference's own file is read by
the same parser (`bin/beam -p <reference>/config/random_positions.txt`;
tests/test_oracle_golden.py::test_reference_random_positions_parse).
"""
import os
import sys
import numpy as np

HALF_FOV = 3.5 * np.pi / 180.0
MAX_POS = 250.0


def _g(v):
    """Python-2 str(float) formatting (12 significant digits), as the notebooks wrote them."""
    s = "%.12g" % v
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s


def _write(path, header, rows, sep):
    with open(path, "w") as f:
        f.write("%d\n" % header)
        f.write("\n".join(sep.join(r) for r in rows))
        f.write("\n")


def linear_positions(n=64):
    x = np.linspace(-MAX_POS, MAX_POS, n)
    return [(repr(float(v)), "0", "0") for v in x]


def grid_positions(n=64):
    k = int(np.floor(np.sqrt(n)))
    x = np.linspace(-MAX_POS, MAX_POS, k)
    xv, yv = np.meshgrid(x, x)
    return [(repr(float(a)), repr(float(b)), "0.0") for a, b in zip(xv.ravel(), yv.ravel())]


def random_positions(n=64, seed=0xD5A):
    rng = np.random.default_rng(seed)
    x = 2 * MAX_POS * rng.uniform(size=n) - MAX_POS
    y = 2 * MAX_POS * rng.uniform(size=n) - MAX_POS
    z = rng.uniform(size=n) * 5 - 2.5
    return [(_g(a), _g(b), _g(c)) for a, b, c in zip(x, y, z)]


def linear_dirs(n, phi_txt):
    th = np.linspace(-HALF_FOV, HALF_FOV, n)
    return [(_g(v), phi_txt) for v in th]


def grid_dirs(n):
    k = int(round(np.sqrt(n)))
    assert k * k == n
    ang = np.linspace(-HALF_FOV, HALF_FOV, k)
    t, p = np.meshgrid(ang, ang)
    return [(_g(a), _g(b)) for a, b in zip(t.ravel(), p.ravel())]


def scaled_dirs(n_theta=64, n_phi=32):
    """BASELINE.json configs[3]: 2048 beam directions, 64 x 32 grid over the same +-3.5 degree field."""
    t, p = np.meshgrid(np.linspace(-HALF_FOV, HALF_FOV, n_theta), np.linspace(-HALF_FOV, HALF_FOV, n_phi))
    return [(_g(a), _g(b)) for a, b in zip(t.ravel(), p.ravel())]


def scaled_sources(n=64, seed=0xD5A):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-HALF_FOV, HALF_FOV, size=(n, 2))
    return [(_g(a), _g(b)) for a, b in ang]


DADA_HEADER = [
    ("HDR_VERSION", "1.0"), ("HDR_SIZE", "4096"), ("INSTRUMENT", "DSAX"), ("TELESCOPE", "DSA-10"),
    ("SOURCE", "TEST"), ("MODE", "RAW"), ("NBIT", "4"), ("NDIM", "1"), ("NPOL", "2"), ("NCHAN", "2048"),
    ("NBEAM", "1"), ("NANT", "2"), ("ANTENNAS", "1-2-5-3"), ("FREQ", "1405.000000"), ("CFREQ", "1405"),
    ("BW", "-250"), ("BANDWIDTH", "-250"), ("TSAMP", "64"), ("RESOLUTION", "4096"), ("ACC_LEN", "1"),
    ("CHAN_AV", "0"), ("FSCRUNCH", "1"), ("TSCRUNCH", "1"), ("DSB", "0"), ("N_PROD", "1"),
    ("OBS_OFFSET", "0"), ("OBS_UNIT", "SECONDS"), ("OBS_VAL", "0000.0000"), ("FILE_SIZE", "2415919104"),
    ("FILE_NUMBER", "0"), ("TRANSFER_SIZE", "126562550000000"), ("UTC_START", "2015-08-07-17:07:28"),
    ("RA", "00:00:00.000"), ("DEC", "00:00:00.000"), ("OBSERVER", "DSA"), ("PID", "P000"), ("RECEIVER", "SANDY"),
]


def make_all(outdir):
    os.makedirs(outdir, exist_ok=True)
    j = lambda n: os.path.join(outdir, n)
    _write(j("linear_positions.txt"), 64, linear_positions(), "\t")
    _write(j("grid_positions.txt"), 64, grid_positions(), " ")
    _write(j("random_positions.txt"), 64, random_positions(), "\t")
    _write(j("linear_directions.txt"), 256, linear_dirs(256, "0"), " ")
    _write(j("grid_beam_directions.txt"), 256, grid_dirs(256), "\t")
    _write(j("linear_source_directions_1024.txt"), 1024, linear_dirs(1024, "0.0"), "\t")
    _write(j("grid_source_directions_3721.txt"), 3721, grid_dirs(3721), "\t")
    _write(j("grid_source_directions_4096.txt"), 4096, grid_dirs(4096), "\t")
    # scaled array (BASELINE.json configs[3]): 16 x 16 antennas over +-250 m, 2048 beams, 64 seeded random sources
    _write(j("scaled_positions_256.txt"), 256, grid_positions(256), " ")
    _write(j("scaled_beam_directions_2048.txt"), 2048, scaled_dirs(), "\t")
    _write(j("scaled_source_directions_64.txt"), 64, scaled_sources(), "\t")
    with open(j("correlator_header_dsaX.txt"), "w") as f:
        for k, v in DADA_HEADER:
            f.write("%-12s %s\n" % (k, v))
    return outdir


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "config")
    print("wrote", make_all(os.path.normpath(out)))
