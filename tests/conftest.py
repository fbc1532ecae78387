"""pytest wiring: marker registration + import paths.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI
symbol presence, gloo world_size-2 sharding.  ``-m gpu`` are the parity tests
proper, on a B200, through the C-ABI.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "glearning-benchmark_b200")
for p in (ROOT, PKG, os.path.join(PKG, "AutoGraph")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_gt():
    return load_golden("ibtt_graph_token.json")


@pytest.fixture(scope="session")
def golden_zinc():
    return load_golden("ibtt_zinc.json")


@pytest.fixture(scope="session")
def golden_agtt():
    return load_golden("agtt_post.json")
