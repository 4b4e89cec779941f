"""pytest configuration: marker registration and shared fixtures.

`-m "not gpu"` (runs in the authoring container, no GPU): oracle vs the reference's golden vectors,
host logic, C-ABI load/symbol checks, gloo world_size-2 sharding logic.
`-m gpu` (runs on a B200): parity tests proper, all through the C-ABI of libzerocaf_b200.so.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5A45524F43414621


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    o.lib()
    return o


def kat_arr(kats, section, name):
    return np.array(kats[section]["items"][name]["values"], dtype=np.uint64)
