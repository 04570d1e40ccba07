import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def shipped40k():
    return load_golden("shipped_40000")


@pytest.fixture(scope="session")
def disk1m_golden():
    return load_golden("disk_1000000")


def golden_inputs(name):
    """(pos, vel, mass, golden dict) of a golden case; prefix cases slice the shipped inputs."""
    g = load_golden(name)
    if "pos" in g:
        return g["pos"], g["vel"], g["mass"], g
    if name.startswith("shipped_"):
        s = load_golden("shipped_40000")
        n = int(g["n"])
        return s["pos"][:n].copy(), s["vel"][:n].copy(), s["mass"][:n].copy(), g
    raise KeyError(name)
