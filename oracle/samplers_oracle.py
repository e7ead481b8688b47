"""CPU ORACLE for the baseline samplers — TEST INFRASTRUCTURE, never imported by the product.

Plain Python / numpy restatement, loop for loop, of
  StaticObjects.collide_sphere            environment/static_objects.py:44-57, :93-106 (+ FCL 0.5.0 sphereSphereIntersect)
  get_random_samples, get_timed_roadmaps_random / _random_common / _fully_random     roadmap/random_sampler.py:22-143
  get_grid, get_timed_roadmaps_grid / _grid_common_2d_fast                           roadmap/grid_sampler.py:19-136
  get_fixed_structured_roadmaps, get_common_roadmaps                                 roadmap/utils.py:45-134
  TimedRoadmap.append_fixed_strucutre / append_samples / add_parents                 roadmap/timed_roadmap.py:116-210
Randomness comes from numpy's global stream exactly as in the reference (np.random.random(dim) per candidate).
Pinned by tests/golden/ref_samplers.npz, written by oracle/gen_golden_samplers.py from the UNMODIFIED reference; the
reference's one FCL query (sphere vs sphere) has no FCL here and is restated: parity UNPINNED at that boundary.
Results: per agent a dict(start, layers=[positions per t], E=[adjacency lists per t]) plus the two counters.
"""
from __future__ import annotations

import math
from itertools import product

import numpy as np

import ctrm_oracle as O


class Counters:
    def __init__(self):
        self.cnt_static_collide = 0


def collide_sphere(oi: O.OInstance, cnt: Counters, pos: np.ndarray, rad: float) -> bool:
    """static_objects.py:44-57, :93-106; the FCL query restated: collide iff not (||c2 - c1|| > r1 + r2)"""
    cnt.cnt_static_collide += 1
    if max(pos + rad) > 1 or min(pos - rad) < 0:
        return True
    for o in range(len(oi.obs_rad)):
        dx, dy = float(oi.obs_pos[o, 0]) - float(pos[0]), float(oi.obs_pos[o, 1]) - float(pos[1])
        if not (math.sqrt(dx * dx + dy * dy + 0.0 * 0.0) > rad + float(oi.obs_rad[o])):
            return True
    return False


def get_random_samples(oi, cnt, num, rad, wo_invalid_samples=True):
    """random_sampler.py:22-51"""
    locs = []
    if wo_invalid_samples:
        while len(locs) < num:
            pos = np.random.random(2) * (1 - 2 * rad) + rad
            if not collide_sphere(oi, cnt, pos, rad):
                locs.append(pos)
    else:
        for _ in range(num):
            pos = np.random.random(2) * (1 - 2 * rad) + rad
            if not collide_sphere(oi, cnt, pos, rad):
                locs.append(pos)
    return locs


def get_grid(oi, cnt, size, rad):
    """grid_sampler.py:19-36"""
    return [np.array(s) for s in product(np.linspace(rad, 1 - rad, size), repeat=2)
            if not collide_sphere(oi, cnt, np.array(s), rad)]


def _fixed(oi, i, locs, T, adjs=None):
    """TimedRoadmap.append_fixed_strucutre (timed_roadmap.py:169-210) for agent i"""
    speed, rad = float(oi.max_speeds[i]), float(oi.rads[i])
    valid = lambda a, b: O.valid_move(oi, a, b, speed, rad)  # noqa: E731
    start = oi.starts[i]
    e0 = [k for k, loc in enumerate(locs) if valid(start, loc)]  # append_sample(loc, 1): add_parents against V[0]
    if adjs is None:
        adjs = [[k] for k in range(len(locs))]
        for a in range(len(locs)):
            for b in range(a + 1, len(locs)):
                if valid(locs[a], locs[b]):
                    adjs[a].append(b)
                    adjs[b].append(a)
    E = [[e0]] + [adjs for _ in range(1, T)] + [[[] for _ in range(len(locs))]]
    return dict(start=start, layers=[np.array(locs).reshape(-1, 2)] * T, E=E)


def fixed_structured(oi, T, arr_locs):
    """roadmap/utils.py:45-84"""
    return [_fixed(oi, i, arr_locs[i], T) for i in range(oi.num_agents)]


def common(oi, T, locs_common, adjs_common=None):
    """roadmap/utils.py:87-134"""
    speed, rad = float(oi.max_speeds[0]), float(oi.rads[0])
    valid = lambda a, b: O.valid_move(oi, a, b, speed, rad)  # noqa: E731
    n = len(locs_common)
    if adjs_common is None:
        adjs_common = [[k] for k in range(n)]
        for a in range(n):
            for b in range(a + 1, n):
                if valid(locs_common[a], locs_common[b]):
                    adjs_common[a].append(b)
                    adjs_common[b].append(a)
    out = []
    for i in range(oi.num_agents):
        locs = list(locs_common) + [oi.goals[i]]
        adjs = [list(a) for a in adjs_common] + [[]]
        for j, loc in enumerate(locs_common):
            if valid(loc, oi.goals[i]):
                adjs[j].append(n)
                adjs[-1].append(j)
        out.append(_fixed(oi, i, locs, T, adjs))
    return out


def random(oi, cnt, T, num, wo_invalid_samples=True):
    """random_sampler.py:54-78"""
    return fixed_structured(oi, T, [get_random_samples(oi, cnt, num, float(oi.rads[i]), wo_invalid_samples) + [oi.goals[i]]
                                    for i in range(oi.num_agents)])


def random_common(oi, cnt, T, num, wo_invalid_samples=True):
    """random_sampler.py:126-143"""
    return common(oi, T, get_random_samples(oi, cnt, num, float(oi.rads[0]), wo_invalid_samples))


def grid(oi, cnt, T, size):
    """grid_sampler.py:39-55"""
    return fixed_structured(oi, T, [get_grid(oi, cnt, size, float(oi.rads[i])) + [oi.goals[i]] for i in range(oi.num_agents)])


def grid_common_2d_fast(oi, cnt, T, size):
    """grid_sampler.py:79-136"""
    rad, speed = float(oi.rads[0]), float(oi.max_speeds[0])
    reach = int(speed / (1 / size))
    valid = lambda a, b: O.valid_move(oi, a, b, speed, rad)  # noqa: E731
    locs, indexes, k = [], np.full((size, size), -1, dtype=np.int32), 0
    for i, s in enumerate(product(np.linspace(rad, 1 - rad, size), repeat=2)):
        if collide_sphere(oi, cnt, np.array(s), rad):
            continue
        indexes[i // size, i % size] = k
        locs.append(np.array(s))
        k += 1
    adjs = [[a] for a in range(len(locs))]
    for i, j in product(range(size), repeat=2):
        idx = indexes[i, j]
        if idx == -1:
            continue
        for _i in range(i, min(i + reach, size - 1) + 1):
            for _j in range(max(0, j - reach), min(j + reach, size - 1) + 1):
                if _i == i and _j == j:
                    continue
                _idx = indexes[_i, _j]
                if _idx < idx:
                    continue
                if valid(locs[idx], locs[_idx]):
                    adjs[idx].append(int(_idx))
                    adjs[_idx].append(int(idx))
    return common(oi, T, locs, adjs)


def fully_random(oi, cnt, T, num, wo_invalid_samples=True):
    """random_sampler.py:81-123: per (t, agent) fresh samples, append_samples -> add_parents against V[t-1]"""
    N = oi.num_agents
    out = [dict(start=oi.starts[i], layers=[], E=[[[]]]) for i in range(N)]
    prev = [[oi.starts[i]] for i in range(N)]
    for t in range(1, T + 1):
        for i in range(N):
            speed, rad = float(oi.max_speeds[i]), float(oi.rads[i])
            locs = get_random_samples(oi, cnt, num, rad, wo_invalid_samples) + [oi.goals[i]]
            Eprev = out[i]["E"][t - 1]
            for k, loc in enumerate(locs):  # append_sample(loc, t, valid_edge): parents in index order
                for p, u in enumerate(prev[i]):
                    if O.valid_move(oi, u, loc, speed, rad):
                        Eprev[p].append(k)
            out[i]["layers"].append(np.array(locs).reshape(-1, 2))
            out[i]["E"].append([[] for _ in locs])
            prev[i] = locs
    return out


def flatten(trms):
    """-> (pos, eptr, eidx) over agents, layers 0..T, vertices in order — the comparison format of the tests"""
    pos, eptr, eidx = [], [0], []
    for tr in trms:
        layers = [np.asarray(tr["start"], dtype=np.float64).reshape(1, 2)] + list(tr["layers"])
        for t, L in enumerate(layers):
            for k in range(len(L)):
                pos.append(L[k])
                eidx.extend(int(e) for e in tr["E"][t][k])
                eptr.append(len(eidx))
    return np.array(pos).reshape(-1, 2), np.array(eptr, dtype=np.int64), np.array(eidx, dtype=np.int32)
