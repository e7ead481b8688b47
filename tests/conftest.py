import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHTS = os.path.join(GOLD, "aamas22_main_best.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def stage_gold():
    return np.load(os.path.join(GOLD, "ref_stage.npz"))


@pytest.fixture(scope="session")
def trace_gold():
    return np.load(os.path.join(GOLD, "ref_trace_demo.npz"))


@pytest.fixture(scope="session")
def oracle_weights():
    import ctrm_oracle as O

    return O.OWeights.load(WEIGHTS)


@pytest.fixture(scope="session")
def builder():
    """one GPU context with the trained networks resident (GPU tests only)"""
    from ctrm_b200 import RoadmapBuilder

    b = RoadmapBuilder(WEIGHTS, 0)
    yield b
    b.close()
