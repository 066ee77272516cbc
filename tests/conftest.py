"""Shared pytest configuration: registers the ``gpu`` marker and puts the repo root, the product
package directory and ``oracle/`` on sys.path (the oracle is test infrastructure only)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ofdm-based-systems_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_link_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "link_*.npz")))


def golden_noisebump_names():
    return sorted(os.path.basename(p)[10:-4] for p in glob.glob(os.path.join(GOLDEN, "noisebump_*.npz")))


def golden_loaded_names():
    return sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "loaded_*.npz")))


def golden_sim_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "sim_*.npz")))


def load_golden(kind, name):
    return np.load(os.path.join(GOLDEN, f"{kind}_{name}.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def kat():
    return np.load(os.path.join(GOLDEN, "kat.npz"), allow_pickle=False)
