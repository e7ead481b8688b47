"""Baseline samplers.  tests/golden/ref_samplers.npz holds runs of the UNMODIFIED reference's
get_timed_roadmaps_random / _fully_random / _random_common / _grid / _grid_common (oracle/gen_golden_samplers.py) on the
instance of the reference's own tests/roadmap/test_random_sampler.py and on a seeded aamas22-shape instance: vertex positions,
adjacency lists IN THE REFERENCE'S ORDER, both collision counters and the state of numpy's global stream after the call.
CPU: the oracle restatement reproduces them.  GPU: ctrm_b200.roadmap (ctrm_collide_points + ctrm_pair_adjacency +
ctrm_adjacency_csr through the C-ABI) reproduces them bit for bit."""
import os

import numpy as np
import pytest

import ctrm_oracle as O
import samplers_oracle as S
from conftest import GOLD

CASES = (  # name, instance, sampler, kwargs, seed  (oracle/gen_golden_samplers.py CASES)
    ("t_random", "test", "random", dict(T=3, num=10), 1),
    ("t_fully", "test", "fully_random", dict(T=3, num=10), 2),
    ("t_rcommon", "test", "random_common", dict(T=3, num=10), 3),
    ("t_grid", "test", "grid", dict(T=3, size=5), 4),
    ("t_gcommon", "test", "grid_common", dict(T=3, size=5), 5),
    ("a_random", "aamas", "random", dict(T=4, num=300), 6),
    ("a_random_inv", "aamas", "random", dict(T=2, num=200, wo_invalid_samples=False), 7),
    ("a_rcommon", "aamas", "random_common", dict(T=3, num=300), 8),
    ("a_grid", "aamas", "grid", dict(T=3, size=24), 9),
    ("a_gcommon", "aamas", "grid_common", dict(T=3, size=40), 10),
    ("a_fully", "aamas", "fully_random", dict(T=3, num=60), 11),
)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ref_samplers.npz"))


def _oins(z, which):
    return O.OInstance.from_arrays(z[f"{which}.starts"], z[f"{which}.goals"], z[f"{which}.max_speeds"], z[f"{which}.rads"],
                                   z[f"{which}.obs_pos"], z[f"{which}.obs_rad"])


@pytest.mark.parametrize("name,which,sampler,kw,seed", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference(gold, name, which, sampler, kw, seed):
    z = gold
    oi = _oins(z, which)
    cnt = S.Counters()
    fn = dict(random=S.random, fully_random=S.fully_random, random_common=S.random_common, grid=S.grid,
              grid_common=S.grid_common_2d_fast)[sampler]
    np.random.seed(seed)
    pos, eptr, eidx = S.flatten(fn(oi, cnt, **kw))
    assert np.array_equal(pos.view(np.uint64), z[f"{name}.pos"].view(np.uint64))
    assert np.array_equal(eptr, z[f"{name}.eptr"]) and np.array_equal(eidx, z[f"{name}.eidx"])
    assert cnt.cnt_static_collide == int(z[f"{name}.cnt_static"])
    assert oi.cnt_continuous_collide == int(z[f"{name}.cnt_continuous"])
    assert np.random.random() == float(z[f"{name}.rng_after"])


def _flatten_product(trms):
    pos, eptr, eidx = [], [0], []
    for trm in trms:
        for t in range(len(trm.V)):
            for i, v in enumerate(trm.V[t]):
                pos.append(np.asarray(v.pos, dtype=np.float64))
                eidx.extend(int(e) for e in trm.E[t][i])
                eptr.append(len(eidx))
    return np.array(pos).reshape(-1, 2), np.array(eptr, dtype=np.int64), np.array(eidx, dtype=np.int32)


@pytest.mark.gpu
@pytest.mark.parametrize("name,which,sampler,kw,seed", CASES, ids=[c[0] for c in CASES])
def test_gpu_samplers_match_reference(gold, name, which, sampler, kw, seed):
    from ctrm_b200 import roadmap as RM
    from helpers import instance_from_oins

    z = gold
    ins = instance_from_oins(_oins(z, which))
    fn = dict(random=RM.get_timed_roadmaps_random, fully_random=RM.get_timed_roadmaps_fully_random,
              random_common=RM.get_timed_roadmaps_random_common, grid=RM.get_timed_roadmaps_grid,
              grid_common=RM.get_timed_roadmaps_grid_common)[sampler]
    np.random.seed(seed)
    trms = fn(ins, **kw)
    assert len(trms) == ins.num_agents  # what the reference's own tests assert (tests/roadmap/test_random_sampler.py:28-32)
    assert np.array_equal(np.array([len(t.V) for t in trms]), z[f"{name}.nlayers"])
    pos, eptr, eidx = _flatten_product(trms)
    assert np.array_equal(pos.view(np.uint64), z[f"{name}.pos"].view(np.uint64))
    assert np.array_equal(eptr, z[f"{name}.eptr"]) and np.array_equal(eidx, z[f"{name}.eidx"])
    assert ins.objs.cnt_static_collide == int(z[f"{name}.cnt_static"])
    assert ins.objs.cnt_continuous_collide == int(z[f"{name}.cnt_continuous"])
    assert np.random.random() == float(z[f"{name}.rng_after"])


@pytest.mark.gpu
def test_gpu_static_point_test_vs_oracle():
    """ctrm_collide_points against the oracle's collide_sphere on random and adversarial (tangent, boundary) points"""
    from ctrm_b200 import roadmap as RM
    from helpers import instance_from_oins

    rng = np.random.RandomState(5)
    O_ = 12
    oi = O.OInstance.from_arrays(np.zeros((1, 2)), np.ones((1, 2)), [0.1], [0.02], rng.rand(O_, 2), rng.rand(O_) * 0.08 + 0.02)
    ins = instance_from_oins(oi)
    rad = 0.03
    pts = rng.rand(4000, 2)
    ang = rng.rand(500) * 2 * np.pi
    k = rng.randint(0, O_, 500)
    pts[:500] = oi.obs_pos[k] + np.stack([np.cos(ang), np.sin(ang)], 1) * (oi.obs_rad[k] + rad)[:, None]  # tangent circles
    pts[500:520, 0] = rad          # touching the unit box from inside
    pts[520:540, 1] = 1 - rad
    got = RM.collide_points(ins, pts, rad)
    cnt = S.Counters()
    want = np.array([S.collide_sphere(oi, cnt, p, rad) for p in pts])
    assert np.array_equal(got, want)
    assert ins.objs.cnt_static_collide == 4000 and 0.2 < want.mean() < 0.9


@pytest.mark.gpu
def test_gpu_pair_adjacency_large_vs_oracle():
    """one set of 1,500 samples (wider than one warp item of 256 columns): edge lists and the counter against the oracle's
    double loop"""
    from ctrm_b200 import roadmap as RM
    from helpers import instance_from_oins

    rng = np.random.RandomState(9)
    oi = O.OInstance.from_arrays(np.zeros((1, 2)) + 0.5, np.ones((1, 2)) * 0.6, [1 / 32], [1 / 64], rng.rand(10, 2),
                                 rng.rand(10) * 0.015 + 0.025)
    ins = instance_from_oins(oi)
    locs = rng.rand(1500, 2) * 0.4 + 0.3
    ep, ei, cnt = RM.pair_adjacency(ins, [locs], [locs], [1 / 32], [1 / 64], symmetric=True, self_first=True)
    adjs = [[k] for k in range(len(locs))]
    for a in range(len(locs)):
        d = np.abs(locs[a + 1:] - locs[a]).max(axis=1)
        for b in (np.nonzero(d <= 1 / 32)[0] + a + 1):  # the L-inf pre-check prunes the oracle's loop, nothing else
            if O.valid_move(oi, locs[a], locs[b], 1 / 32, 1 / 64):
                adjs[a].append(int(b))
                adjs[b].append(int(a))
    assert [ei[ep[r]:ep[r + 1]].tolist() for r in range(len(locs))] == adjs
    assert int(cnt[0]) == oi.cnt_continuous_collide
