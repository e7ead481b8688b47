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


# full recorded runs of the UNMODIFIED reference on the demo instance (oracle/gen_golden.py): N_traj = 3 (= dim_ind, every
# roll-out learned), and two with N_traj > dim_ind where the reference itself decides the gate / p_rw(t, makespan) / dense
# merge rules — eval parameters (rate 0.1) and the function default (rate 0.25)
TRACES = ("ref_trace_demo.npz", "ref_trace_demo_nt8_r010.npz", "ref_trace_demo_nt6_r025.npz")


@pytest.fixture(scope="session", params=TRACES)
def any_trace_gold(request):
    return np.load(os.path.join(GOLD, request.param))


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
