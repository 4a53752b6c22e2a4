import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CONFIG_DIR = os.path.join(ROOT, "config")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def config_dir():
    """config/*.txt are regenerated on demand (tools/make_configs.py)."""
    import make_configs
    if not os.path.exists(os.path.join(CONFIG_DIR, "grid_source_directions_4096.txt")):
        make_configs.make_all(CONFIG_DIR)
    return CONFIG_DIR


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    with open(os.path.join(GOLDEN_DIR, "ref_host_golden.json")) as f:
        g = json.load(f)
    small = dict(np.load(os.path.join(GOLDEN_DIR, "ref_host_small.npz")))
    return g, small


@pytest.fixture(scope="session")
def bf():
    """The product library; GPU tests fail (not skip) if it is not built."""
    import dsabeamformer_b200 as m
    m.lib()
    return m
