"""CPU tests (no GPU) of the host side: the C-ABI library loads and exports every symbol include/ctrm_b200.h declares,
weights pack into the documented blob layout, instances pack into the SoA upload layout, the CSR -> TimedRoadmap views
behave like the reference container, and the product path fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import pickle
import re

import numpy as np
import pytest

from conftest import ROOT, WEIGHTS


def test_library_exports_every_declared_symbol():
    from ctrm_b200 import _lib

    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "ctrm_b200.h")).read()
    declared = set(re.findall(r"\b(ctrm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ctrm_status"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libctrm_b200.so does not export {name}"
    assert set(_lib.SIGNATURES) == declared, (set(_lib.SIGNATURES) ^ declared)
    assert lib.ctrm_abi_version() == _lib.ABI_VERSION


def test_ctypes_signatures_have_the_declared_arity():
    """the Python binding (ctrm_b200/_lib.py) passes exactly as many arguments as include/ctrm_b200.h declares for every
    entry point: a parameter added on one side only would otherwise shift every later argument silently"""
    from ctrm_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "ctrm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)      # comments carry commas and parentheses
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(", hdr)
        assert m, name
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(hdr[i], 0)
            i += 1
        params = hdr[m.end():i - 1].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), f"{name}: header declares {n} parameters, _lib.SIGNATURES passes {len(argtypes)}"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ctrm_b200 import RoadmapBuilder, _lib

    with pytest.raises(_lib.CtrmError):
        RoadmapBuilder(WEIGHTS, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ctrm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "ctrm_oracle" not in src and "import oracle" not in src and "liboracle" not in src, f


def test_weight_blob_layout():
    from ctrm_b200 import weights

    b = weights.load_bundle(WEIGHTS)
    desc, off, blob = weights.pack_blob(b)
    assert blob.dtype == np.float32 and blob.size == off.total
    a = b.arrays
    KD = 2 * b.fov_size ** 2
    w0 = blob[off.fov_w0:off.fov_w0 + KD * 64].reshape(KD, 64)
    assert np.array_equal(w0[:, :32], a["fov_self.mlp.0.weight"].T) and np.array_equal(w0[:, 32:], a["fov_vain.mlp.0.weight"].T)
    ag = a["agent.mlp.0.weight"].T
    assert np.array_equal(blob[off.ag_w0i:off.ag_w0i + 11 * 32].reshape(11, 32), ag[:11])
    assert np.array_equal(blob[off.ag_w0v:off.ag_w0v + 32 * 32].reshape(32, 32), ag[11:])
    w2 = blob[off.ag_w2:off.ag_w2 + 32 * 44].reshape(32, 44)
    assert np.array_equal(w2[:, :42], a["agent.mlp.4.weight"].T) and not w2[:, 42:].any()
    # eval-mode BatchNorm folded into the Linear: y = BN(Wx + b) == W'x + b'
    rng = np.random.RandomState(0)
    x = rng.randn(5, 139)
    W, bb = a["model.decoder.0.weight"].astype(np.float64), a["model.decoder.0.bias"].astype(np.float64)
    g, be = a["model.decoder.1.weight"].astype(np.float64), a["model.decoder.1.bias"].astype(np.float64)
    mu, var = a["model.decoder.1.running_mean"].astype(np.float64), a["model.decoder.1.running_var"].astype(np.float64)
    ref = (x @ W.T + bb - mu) / np.sqrt(var + 1e-5) * g + be
    Wf = blob[off.dec_w0:off.dec_w0 + 139 * 32].reshape(139, 32).astype(np.float64)
    bf = blob[off.dec_b0:off.dec_b0 + 32].astype(np.float64)
    assert np.abs(x @ Wf + bf - ref).max() < 1e-5
    for name, _ in off._fields_:
        assert getattr(off, name) % 4 == 0  # 16-byte aligned blocks


def test_weights_from_reference_files_match_npz():
    ref = "/root/reference/trained_models/aamas22-main/best"
    if not os.path.exists(ref + "_model.pt"):
        pytest.skip("reference checkout not present (GPU box)")
    from ctrm_b200 import weights

    b1, b2 = weights.load_bundle(WEIGHTS), weights.load_bundle(ref)
    assert np.array_equal(weights.pack_blob(b1)[2], weights.pack_blob(b2)[2])
    assert (b2.fov_size, b2.map_size, b2.num_neighbors, b2.dim_ind, b2.temp) == (19, 160, 15, 3, 2.0)
    assert weights.resolve_basename("2021-10-01_12-00-00") == "/workspace/log/2021-10-01_12-00-00/best"


def test_instance_mirror_and_pickle(trace_gold):
    from ctrm_b200 import Instance, ObstacleSphere, RoadmapBuilder
    from helpers import instance_from_oins, oins_from_npz

    ins = instance_from_oins(oins_from_npz(trace_gold))
    assert ins.num_agents == 28 and len(ins.obs) == 10 and ins.dim == 2
    ins2 = pickle.loads(pickle.dumps(ins))  # objs dropped on pickling and rebuilt (instance.py:37-46)
    assert ins2.objs is not ins.objs and len(ins2.objs.obs) == 10 and ins2.objs.cnt_continuous_collide == 0
    with pytest.raises(Exception):
        ins.num_agents = 3  # frozen
    other = Instance(num_agents=1, starts=[np.array([0.1, 0.1])], goals=[np.array([0.9, 0.9])], max_speeds=[0.05], rads=[0.02],
                     goal_rads=[0.02], obs=[ObstacleSphere(np.array([0.5, 0.5]), 0.1)])
    p = RoadmapBuilder.pack_instances([ins, other])
    assert p["n_agents"].tolist() == [28, 1] and p["n_obs"].tolist() == [10, 1]
    assert p["starts"].shape == (29, 2) and p["obs"].shape == (11, 3) and p["speeds2"][-1] == 0.05 ** 2


def test_reference_instance_pickle_loads():
    path = "/root/reference/notebooks/CTRM_demo_ins.pkl"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    from ctrm_b200 import load_instance

    ins = load_instance(path)
    assert ins.num_agents == 28 and len(ins.obs) == 10 and abs(ins.rads[0] - 1 / 64) < 1e-12


def test_csr_views_and_timed_roadmap_container():
    from ctrm_b200.timed_roadmap import CSRRoadmaps, TimedNode, TimedRoadmap

    # one instance, one agent, L = 4 layer slots, 3 layers used: V = [[s], [a, g], [g]], E = [[0, 1], [[0], [0]], [[]]]
    layer_ptr = np.array([0, 1, 3, 4, 4], dtype=np.int32)
    pos = np.array([[0.1, 0.1], [0.2, 0.1], [0.9, 0.9], [0.9, 0.9]])
    eptr = np.array([0, 2, 3, 4, 4], dtype=np.int64)
    eidx = np.array([0, 1, 0, 0], dtype=np.int32)
    csr = CSRRoadmaps(np.array([0, 1], dtype=np.int32), 4, layer_ptr, pos, eptr, eidx, np.array([3], dtype=np.int32),
                      np.array([7], dtype=np.int64))
    assert csr.instance_counts(0) == (4, 4)
    for lazy in (True, False):
        trm = csr.timed_roadmaps(0, lazy=lazy)[0]
        assert len(trm.V) == 3 and trm.E == [[[0, 1]], [[0], [0]], [[]]]
        assert trm.V[1][1] == TimedNode(1, 1, np.array([0.9, 0.9]))
        assert trm.get_parents(trm.V[2][0]) == [0, 1]
        assert trm.getNodesPosOnestep(1).shape == (2, 2) and trm.getNodesPosOnestep(9).shape == (0, 2)
    # the reference's container test (tests/roadmap/test_timed_roadmap.py:6-19)
    t2 = TimedRoadmap(np.array([0.0, 0.0]))
    valid = lambda a, b: np.linalg.norm(a - b) <= 0.2  # noqa: E731
    t2.append_samples([np.array([0.1, 0.0]), np.array([0.0, 0.1]), np.array([0.3, 0.0])], 1, valid)
    assert len(t2.V[1]) == 3 and t2.E[0][0] == [0, 1]
    t2.append_sample(np.array([0.2, 0.0]), 2, valid)
    assert t2.E[1] == [[0], [], [0]]


def test_prob_random_walk_table_matches_reference_expression():
    from ctrm_b200.roadmap_learned import prob_random_walk_table

    tab = prob_random_walk_table(64, 5.0, 0.0)
    for t, D in ((1, 64), (64, 64), (17, 33), (5, 1)):
        assert tab[t, D] == (1 - np.exp(-5.0 * t / D)) * (1 - 0.0) + 0.0


def test_ctrm_alias_package_resolves_and_unpickles():
    """ctrm_b200/shims on sys.path makes `ctrm.*` of the replaced paths importable (hydra _target_ strings, reference pickles)"""
    import subprocess
    import sys

    code = (
        "import sys, pickle, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import ctrm.roadmap_learned as rl, ctrm.environment as env, ctrm.roadmap as rm, ctrm.planner as pl\n"
        "import ctrm_b200\n"
        "assert rl.get_timed_roadmaps_multiple_paths_with_learned_indicator is ctrm_b200.get_timed_roadmaps_multiple_paths_with_learned_indicator\n"
        "assert env.Instance is ctrm_b200.Instance and pl.PrioritizedPlanning is ctrm_b200.PrioritizedPlanning\n"
        "assert callable(rm.get_timed_roadmaps_random) and rm.TimedRoadmap is ctrm_b200.TimedRoadmap\n"
        "p = '/root/reference/notebooks/CTRM_demo_ins.pkl'\n"
        "if os.path.exists(p):\n"
        "    ins = pickle.load(open(p, 'rb'))\n"
        "    assert type(ins) is ctrm_b200.Instance and ins.num_agents == 28 and len(ins.objs.obs) == 10\n"
        "print('ok')\n") % (os.path.join(ROOT, "ctrm_b200", "shims"), ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
