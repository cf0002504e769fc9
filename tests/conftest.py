"""pytest configuration: the `gpu` marker, repo root on sys.path, golden-fixture loader."""
import os

# several engines (emulated ranks) of one process spin on each other's flags: each of their streams needs its own hardware queue,
# or a waiting kernel sits in front of the very kernels it waits for (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def subtree(d, prefix):
    """{'a/b/c': v} -> {'c': v} for keys starting with 'a/b/'."""
    p = prefix.rstrip("/") + "/"
    return {k[len(p):]: v for k, v in d.items() if k.startswith(p)}


@pytest.fixture(scope="session")
def golden():
    return load_golden
