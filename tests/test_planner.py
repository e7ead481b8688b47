"""Prioritized planning over timed roadmaps.

CPU: the oracle restatement (oracle/pp_oracle.py) against the fixture written by the UNMODIFIED reference planner
(oracle/gen_golden_pp.py): an unsolvable case on the reference's own roadmaps and a solved case on N_traj = 25 roadmaps.
GPU: ctrm_plan_prioritized against the same fixture and, on synthetic batches, against the oracle — chosen vertices,
solved flags and the planner's counters (collision checks, expanded, explored) must be identical."""
import os

import numpy as np
import pytest

import pp_oracle as P
from conftest import GOLD as GOLDEN, WEIGHTS

KEYS = ("nlayers", "layer_ptr", "pos", "eptr", "eidx")


@pytest.fixture(scope="module")
def pp_gold():
    return np.load(os.path.join(GOLDEN, "ref_pp_demo.npz"))


@pytest.fixture(scope="module")
def demo():
    return np.load(os.path.join(GOLDEN, "ref_trace_demo.npz"))


def _arrays(z, prefix):
    return tuple(z[prefix + k] for k in KEYS)


def _check(got, z, tag):
    assert bool(got["solved"]) == bool(z[tag + ".solved"])
    for i, k in enumerate(("cnt_continuous_collide", "lowlevel_expanded", "lowlevel_explored")):
        assert int(got[k]) == int(z[f"{tag}.{k}"]), k
    if bool(z[tag + ".solved"]):
        assert np.array_equal(np.asarray(got["paths"]), z[tag + ".paths"])


def test_oracle_matches_reference_planner_unsolved(pp_gold, demo):
    got = P.prioritized_planning(demo["starts"], demo["goals"], demo["max_speeds"], demo["rads"], P.FlatRoadmaps(*_arrays(demo, "trm.")))
    _check(got, pp_gold, "a")
    assert got["failed_agent"] >= 0 and got["paths"] is None


def test_oracle_matches_reference_planner_solved(pp_gold, demo):
    trms = P.FlatRoadmaps(*_arrays(pp_gold, "b.trm."))
    got = P.prioritized_planning(demo["starts"], demo["goals"], demo["max_speeds"], demo["rads"], trms)
    _check(got, pp_gold, "b")
    pos = P.path_positions(trms, got["paths"])
    costs = [P.get_cost(pos[a], demo["goals"][a], float(pp_gold["goal_rads"][a])) for a in range(len(pos))]
    assert sum(costs) == int(pp_gold["b.sum_of_costs"]) and max(costs) == int(pp_gold["b.maximum_costs"])
    assert abs(sum(P.get_travel_dist(q) for q in pos) - float(pp_gold["b.sum_of_travel_dists"])) < 1e-9


def _strided(arrays):
    """oracle/golden layout (layer_ptr over existing layers only) -> the C-ABI's fixed stride of L slots per agent"""
    nlayers, layer_ptr, pos, eptr, eidx = arrays
    L = int(np.max(nlayers)) + 1
    lp = np.zeros(len(nlayers) * L + 1, dtype=np.int32)
    k = 0
    for a, n in enumerate(nlayers):
        n = int(n)
        lp[a * L:a * L + n + 1] = layer_ptr[k:k + n + 1]
        lp[a * L + n + 1:(a + 1) * L + 1] = layer_ptr[k + n]
        k += n
    return L, lp


def _gpu_solve(arrays, starts, goals, speeds, rads, agent_off=None):
    from ctrm_b200.planner import plan_prioritized_arrays

    L, lp = _strided(arrays)
    N = len(arrays[0])
    off = np.array([0, N], dtype=np.int32) if agent_off is None else agent_off
    return plan_prioritized_arrays(off, L, arrays[0], lp, arrays[2], arrays[3], arrays[4], goals, speeds, rads), L


@pytest.mark.gpu
@pytest.mark.parametrize("tag,src", [("a", "demo"), ("b", "gold")])
def test_gpu_planner_matches_reference(pp_gold, demo, tag, src):
    arrays = _arrays(demo, "trm.") if src == "demo" else _arrays(pp_gold, "b.trm.")
    r, L = _gpu_solve(arrays, demo["starts"], demo["goals"], demo["max_speeds"], demo["rads"])
    T = int(arrays[0][0]) - 1
    got = dict(solved=r["solved"][0], cnt_continuous_collide=r["counters"][0, 0], lowlevel_expanded=r["counters"][0, 1],
               lowlevel_explored=r["counters"][0, 2], paths=r["paths"][:, :T + 1])
    _check(got, pp_gold, tag)
    if not got["solved"]:
        assert r["failed_agent"][0] >= 0 and np.all(r["paths"] == -1)


@pytest.mark.gpu
def test_gpu_planner_batch_vs_oracle_and_drop_in_class():
    """roadmaps built by the GPU for a mixed batch, planned on the device, checked instance by instance against the oracle;
    then the reference-shaped class on one instance"""
    from ctrm_b200 import PrioritizedPlanning, RoadmapBuilder
    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen
    from helpers import csr_to_arrays

    specs = [(71, 6, 9, 4), (72, 10, 14, 8), (73, 3, 4, 0), (74, 18, 20, 10)]
    instances = [gen(lo, hi, "0.03125", "0.015625", o, 0.05, 0.08, rng=np.random.RandomState(s)) for s, lo, hi, o in specs]
    bld = RoadmapBuilder(WEIGHTS, 0)
    try:
        csr = bld.build(instances, seed=9, N_traj=8, max_T=48, max_attempt=3, merge_distance_rate=0.1).copy()
        r = bld.plan()
        paths, solved, counters = r["paths"].cpu().numpy(), r["solved"].cpu().numpy(), r["counters"].cpu().numpy()
        a0 = 0
        for b, ins in enumerate(instances):
            starts, goals, speeds, rads, _ = ins.arrays()
            N = ins.num_agents
            want = P.prioritized_planning(starts, goals, speeds, rads, P.FlatRoadmaps(*csr_to_arrays(csr, b)))
            assert bool(solved[b]) == want["solved"], b
            assert [int(x) for x in counters[b]] == [want["cnt_continuous_collide"], want["lowlevel_expanded"],
                                                      want["lowlevel_explored"]], b
            if want["solved"]:
                T = want["paths"].shape[1] - 1
                assert np.array_equal(paths[a0:a0 + N, :T + 1], want["paths"]), b
            a0 += N
        assert solved.sum() >= 2
        # the reference-shaped class, fed with the lazy TimedRoadmap views of instance 1
        b = int(np.nonzero(solved)[0][0])
        res = PrioritizedPlanning(instances[b], csr.timed_roadmaps(b)).solve()
        assert res.solved and len(res.paths) == instances[b].num_agents
        for i, path in enumerate(res.paths):
            assert np.array_equal(path[0].pos, instances[b].starts[i]) and np.array_equal(path[-1].pos, instances[b].goals[i])
            d = np.diff(np.stack([v.pos for v in path]), axis=0)
            assert np.all(np.linalg.norm(d, axis=1) <= instances[b].max_speeds[i] * (1 + 1e-12))
        assert res.lowlevel_explored == int(counters[b][2]) and res.sum_of_costs > 0
        # solve() = _solve() + validate() (planner/planner.py:103-105): a valid solution adds one swept test per agent pair
        # and time step to the counter, exactly as the reference's validate does through collide_dynamic_agents
        N_b, T_b = instances[b].num_agents, len(res.paths[0]) - 1
        assert res.cnt_continuous_collide == int(counters[b][0]) + T_b * N_b * (N_b - 1) // 2
    finally:
        bld.close()


@pytest.mark.gpu
def test_demonstration_features_vs_oracle():
    """FOV / others_info / self_info of a planned demonstration == the oracle's, bit for bit"""
    import ctrm_oracle as O
    from ctrm_b200 import PrioritizedPlanning, RoadmapBuilder
    from ctrm_b200.dataset import demonstration_features
    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen

    ins = gen(7, 9, "0.03125,0.046875", "0.015625,0.0234375", 6, 0.05, 0.08, rng=np.random.RandomState(123))
    bld = RoadmapBuilder(WEIGHTS, 0)
    try:
        csr = bld.build([ins], seed=4, N_traj=10, max_T=48, max_attempt=3, merge_distance_rate=0.1).copy()
    finally:
        bld.close()
    pp = PrioritizedPlanning(ins, csr.timed_roadmaps(0))
    res = pp.solve()
    assert res.solved
    T = int(res.maximum_costs)
    got = demonstration_features(ins, pp.path_pos, T=T)
    starts, goals, speeds, rads, obs = ins.arrays()
    N, M, F = ins.num_agents, 160, 19
    occ = O.draw_occupancy(obs[:, :2], obs[:, 2], M)
    ctgs = np.stack([O.cost_to_go(goals[i], occ, float(rads[i])) for i in range(N)])
    for t in range(T):
        cur = pp.path_pos[:, t]
        prev = pp.path_pos[:, max(0, t - 1)]
        assert np.array_equal(got["fovs"][t], O.fov_crop_batch(cur, occ, ctgs, F).astype(np.float32)), t
        want = O.others_info(cur, goals, prev, rads, speeds).astype(np.float32)
        assert np.array_equal(got["others_info"][t], want), t
        assert np.array_equal(got["self_info"][t], want.reshape(N, N, 11)[np.arange(N), np.arange(N), 3:]), t


def test_flatten_trms_lazy_views_equal_plain_lists(pp_gold):
    """host logic of the planner mirror (no GPU): the CSR fast path over lazy views and the generic path over plain
    TimedRoadmap lists describe the same layered graph"""
    from ctrm_b200.planner import _flatten_trms
    from ctrm_b200.timed_roadmap import CSRRoadmaps

    nlayers, layer_ptr, pos, eptr, eidx = _arrays(pp_gold, "b.trm.")
    N = 6  # the first six agents are enough
    L, lp = _strided((nlayers[:N], layer_ptr, pos, eptr, eidx))
    nv = int(lp[-1])
    csr = CSRRoadmaps(np.array([0, N], dtype=np.int32), L, lp, pos[:nv], eptr[:nv + 1], eidx[:int(eptr[nv])],
                      nlayers[:N].astype(np.int32), np.zeros(1, dtype=np.int64))
    lazy = csr.timed_roadmaps(0, lazy=True)
    plain = csr.timed_roadmaps(0, lazy=False)
    a = _flatten_trms(lazy)
    b = _flatten_trms(plain)
    assert a[0] >= int(nlayers[0]) + 1 and b[0] == int(nlayers[0]) + 1
    assert np.array_equal(a[1], b[1])                      # n_layers
    assert np.array_equal(a[3], b[3])                      # positions
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])   # CSR adjacency
    for ag in range(N):                                    # layer starts agree wherever a layer exists
        n = int(a[1][ag])
        assert np.array_equal(a[2][ag * a[0]:ag * a[0] + n + 1], b[2][ag * b[0]:ag * b[0] + n + 1])


def _random_layered_roadmaps(rng, N, T, width):
    """small random timed roadmaps in the flat layout: V[0] = [start], V[t][-1] = goal, random forward edges"""
    starts, goals = rng.rand(N, 2), rng.rand(N, 2)
    nlayers, layer_ptr, pos, eptr, eidx = [], [0], [], [0], []
    for a in range(N):
        layers = [[starts[a]]] + [[rng.rand(2) for _ in range(width - 1)] + [goals[a]] for _ in range(T)]
        nlayers.append(len(layers))
        for t, layer in enumerate(layers):
            for _ in layer:
                nxt = len(layers[t + 1]) if t + 1 < len(layers) else 0
                adj = sorted(set(rng.randint(0, nxt, size=rng.randint(1, 4)).tolist())) if nxt else []
                eidx.extend(adj)
                eptr.append(len(eidx))
            pos.extend(layer)
            layer_ptr.append(len(pos))
    return starts, goals, (np.array(nlayers), np.array(layer_ptr), np.array(pos), np.array(eptr), np.array(eidx, dtype=np.int64))


@pytest.mark.parametrize("seed", [0, 3, 6, 7])  # 3 and 7 are solved, 0 and 6 fail at a late agent
def test_oracle_planner_solutions_are_valid(seed):
    """self-consistency of the planner restatement on random layered graphs: every returned path follows edges from the
    start to the goal vertex, waits at the goal afterwards, and no two agents collide on any step"""
    rng = np.random.RandomState(seed)
    N, T, width = 5, 9, 6
    starts, goals, arrays = _random_layered_roadmaps(rng, N, T, width)
    speeds, rads = np.full(N, 10.0), np.full(N, 0.03)
    trms = P.FlatRoadmaps(*arrays)
    got = P.prioritized_planning(starts, goals, speeds, rads, trms)
    assert got["lowlevel_explored"] >= 1
    if not got["solved"]:
        assert 0 <= got["failed_agent"] < N
        return
    paths = got["paths"]
    assert paths.shape == (N, T + 1)
    xy = P.path_positions(trms, paths)
    for a in range(N):
        assert paths[a, 0] == 0 and np.array_equal(xy[a, 0], starts[a])
        assert np.array_equal(xy[a, -1], goals[a]) and paths[a, -1] == trms.width(a, T) - 1
        arrived = False
        for t in range(T):
            if paths[a, t] == trms.width(a, t) - 1 and t >= 1:
                arrived = arrived or all(paths[a, u] == trms.width(a, u) - 1 for u in range(t, T + 1))
            if not arrived:
                assert paths[a, t + 1] in trms.succ(a, t, int(paths[a, t])), (a, t)
    for t in range(T):
        for i in range(N):
            for j in range(i):
                assert not P.sweep_pair(xy[i, t], xy[i, t + 1], rads[i], xy[j, t], xy[j, t + 1], rads[j]), (t, i, j)


def _batch_layout(insts, L):
    """several instances' flat roadmaps -> one batch in the strided layout of the C-ABI"""
    n_layers, lps, poss, eptrs, eidxs, goals_all, off = [], [], [], [], [], [], [0]
    v_base = e_base = 0
    for starts, goals, arrays in insts:
        nl, lp_flat, pos, eptr, eidx = arrays
        Li, lp = _strided((nl, lp_flat, pos, eptr, eidx))
        assert Li == L
        lps.append(lp[:-1].astype(np.int64) + v_base)
        n_layers.append(nl)
        poss.append(pos)
        eptrs.append(eptr[1:].astype(np.int64) + e_base)
        eidxs.append(eidx)
        goals_all.append(goals)
        v_base += len(pos)
        e_base += len(eidx)
        off.append(off[-1] + len(nl))
    return (np.array(off, dtype=np.int32), np.concatenate(n_layers).astype(np.int32),
            np.concatenate(lps + [np.array([v_base])]).astype(np.int32), np.concatenate(poss),
            np.concatenate([np.zeros(1, dtype=np.int64)] + eptrs), np.concatenate(eidxs).astype(np.int32), np.concatenate(goals_all))


def _random_insts():
    insts = []
    for seed in (0, 3, 6, 7, 11):
        rng = np.random.RandomState(seed)
        insts.append(_random_layered_roadmaps(rng, 3 + seed % 4, 9, 6))
    return insts


def test_batch_layout_round_trip():
    """host side of the batched planner call (no GPU): slicing the concatenated strided layout gives each instance back"""
    insts = _random_insts()
    off, n_layers, lp, pos, eptr, eidx, goals = _batch_layout(insts, 11)
    assert lp[-1] == len(pos) and eptr[-1] == len(eidx) and len(eptr) == len(pos) + 1
    for b, (starts, g, arrays) in enumerate(insts):
        a0, a1 = int(off[b]), int(off[b + 1])
        v0, v1 = int(lp[a0 * 11]), int(lp[a1 * 11])
        assert np.array_equal(pos[v0:v1], arrays[2]) and np.array_equal(goals[a0:a1], g)
        assert np.array_equal(eptr[v0:v1 + 1] - eptr[v0], arrays[3]) and np.array_equal(eidx[eptr[v0]:eptr[v1]], arrays[4])
        for a in range(a0, a1):
            n = int(n_layers[a])
            assert n == 10 and lp[a * 11] - v0 == arrays[1][(a - a0) * 10]
            assert lp[a * 11 + n] == (lp[(a + 1) * 11] if a + 1 < len(n_layers) else len(pos))


@pytest.mark.gpu
def test_gpu_planner_random_layered_graphs_vs_oracle():
    """foreign roadmaps (not built by this package): sparse random edges, several instances of different sizes per call"""
    from ctrm_b200.planner import plan_prioritized_arrays

    insts = _random_insts()
    off, n_layers, lp, pos, eptr, eidx, goals = _batch_layout(insts, 11)
    A = int(off[-1])
    r = plan_prioritized_arrays(off, 11, n_layers, lp, pos, eptr, eidx, goals, np.full(A, 10.0), np.full(A, 0.03))
    for b, (starts, g, arrays) in enumerate(insts):
        N = len(starts)
        want = P.prioritized_planning(starts, g, np.full(N, 10.0), np.full(N, 0.03), P.FlatRoadmaps(*arrays))
        assert bool(r["solved"][b]) == want["solved"], b
        assert [int(x) for x in r["counters"][b]] == [want["cnt_continuous_collide"], want["lowlevel_expanded"],
                                                      want["lowlevel_explored"]], b
        assert int(r["failed_agent"][b]) == want["failed_agent"], b
        if want["solved"]:
            assert np.array_equal(r["paths"][off[b]:off[b + 1], :10], want["paths"]), b
