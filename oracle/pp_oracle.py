"""CPU restatement of the reference's prioritized planning over timed roadmaps — TEST INFRASTRUCTURE ONLY.

Nothing under ctrm_b200/ imports this file; it is the checker for the GPU planner (tests/, smoke(), bench cpu legs).

Follows, function by function:
  PrioritizedPlanning._solve            src/ctrm/planner/prioritized_planning.py:40-101
  Planner.get_single_agent_path         src/ctrm/planner/planner.py:311-384   (space-time A*, CLOSE set on generation)
  Planner.collide_dynamic_agents        src/ctrm/planner/planner.py:216-263   -> continuousCollideSpheres
                                        (cpp_helper/sphere_collision_check_wrapper/src/collision_check.cpp:16-58)
  get_cost / get_travel_dist            src/ctrm/planner/utils.py:13-54

Roadmaps come as the flat arrays the GPU exports (nlayers, layer_ptr, pos, eptr, eidx) for ONE instance.
Tie-breaking: the reference orders OPEN by (f, t, np.random.rand()); exact (f, t) ties need two vertices of one layer at the
same distance from the goal and are broken here by vertex index instead of a random number (the GPU does the same).
Pinned against the reference run in the build container: tests/golden/ref_pp_*.npz (oracle/gen_golden_pp.py).
"""
from __future__ import annotations

import heapq
import math

import numpy as np


def sweep_pair(f1, t1, r1, f2, t2, r2) -> bool:
    """continuousCollideSpheres, both spheres moving (collision_check.cpp:16-58), operation order preserved"""
    a = 0.0
    b = 0.0
    for i in range(2):
        val = ((-f1[i] + t1[i]) + f2[i]) - t2[i]
        a += val * val
        b += val * (f1[i] - f2[i])
    if b >= 0.0:
        tt = 0.0
    elif a + b <= 0.0:
        tt = 1.0
    else:
        tt = (-b) / a
    tot = 0.0
    for i in range(2):
        val = ((1.0 - tt) * f1[i] + tt * t1[i]) - ((1.0 - tt) * f2[i] + tt * t2[i])
        tot += val * val
    rs = r1 + r2
    return tot <= rs * rs


class FlatRoadmaps:
    """one instance's timed roadmaps as flat arrays: vertex (agent, t, i) -> global id, successors as local indices"""

    def __init__(self, nlayers, layer_ptr, pos, eptr, eidx):
        self.nlayers = np.asarray(nlayers, dtype=np.int64)
        self.pos = np.asarray(pos, dtype=np.float64)
        self.eptr = np.asarray(eptr, dtype=np.int64)
        self.eidx = np.asarray(eidx, dtype=np.int64)
        lp = np.asarray(layer_ptr, dtype=np.int64)
        self.first = []  # first[agent][t] = global id of vertex (t, 0); first[agent][nlayers] = end
        k = 0
        for n in self.nlayers:
            self.first.append(lp[k:k + int(n) + 1].copy())
            k += int(n)
        assert k + 1 == len(lp), (k, len(lp))

    def width(self, agent, t):
        return int(self.first[agent][t + 1] - self.first[agent][t])

    def vid(self, agent, t, i):
        return int(self.first[agent][t]) + i

    def succ(self, agent, t, i):
        v = self.vid(agent, t, i)
        return self.eidx[self.eptr[v]:self.eptr[v + 1]]


def prioritized_planning(starts, goals, max_speeds, rads, trms: FlatRoadmaps):
    """-> dict(solved, paths [N][T+1] vertex index per time step (or None), cnt_continuous_collide, lowlevel_expanded,
    lowlevel_explored, failed_agent)"""
    N = len(starts)
    T = int(trms.nlayers[0]) - 1
    required = 1
    solution = []  # per agent: list of local indices, one per t
    cnt = expanded = explored = 0
    for agent in range(N):
        goal, speed, rad = goals[agent], float(max_speeds[agent]), float(rads[agent])

        def fval(t, i):
            p = trms.pos[trms.vid(agent, t, i)]
            d0, d1 = goal[0] - p[0], goal[1] - p[1]
            return t + math.sqrt(d0 * d0 + d1 * d1) / speed

        closed = [np.zeros(trms.width(agent, t), dtype=bool) for t in range(T + 1)]
        parent = [np.full(trms.width(agent, t), -1, dtype=np.int64) for t in range(T + 1)]
        OPEN = [(fval(0, 0), 0, 0)]
        expanded += 1
        path = None
        while OPEN:
            _, t, i = heapq.heappop(OPEN)
            explored += 1
            if t >= required and i == trms.width(agent, t) - 1:  # check_fin: the last vertex of a layer is the goal
                path = [i]
                while t > 0:
                    i = int(parent[t][i])
                    t -= 1
                    path.append(i)
                path.reverse()
                break
            if t == T:
                continue
            pv = trms.pos[trms.vid(agent, t, i)]
            for j in trms.succ(agent, t, i):
                j = int(j)
                if closed[t + 1][j]:
                    continue
                pu = trms.pos[trms.vid(agent, t + 1, j)]
                hit = False
                for k in range(agent):  # the reference builds the whole list, then any(): every pair is counted
                    cnt += 1
                    qa = trms.pos[trms.vid(k, t, solution[k][t])]
                    qb = trms.pos[trms.vid(k, t + 1, solution[k][t + 1])]
                    hit = sweep_pair(pv, pu, rad, qa, qb, float(rads[k])) or hit
                if hit:
                    continue
                closed[t + 1][j] = True
                parent[t + 1][j] = i
                heapq.heappush(OPEN, (fval(t + 1, j), t + 1, j))
                expanded += 1
        if path is None:
            return dict(solved=False, paths=None, cnt_continuous_collide=cnt, lowlevel_expanded=expanded,
                        lowlevel_explored=explored, failed_agent=agent)
        required = max(len(path) - 1, required)
        path += [trms.width(agent, t) - 1 for t in range(len(path), T + 1)]  # wait at the goal
        solution.append(path)
    return dict(solved=True, paths=np.array(solution, dtype=np.int32), cnt_continuous_collide=cnt,
                lowlevel_expanded=expanded, lowlevel_explored=explored, failed_agent=-1)


def path_positions(trms: FlatRoadmaps, paths):
    return np.stack([np.stack([trms.pos[trms.vid(a, t, int(i))] for t, i in enumerate(p)]) for a, p in enumerate(paths)])


def get_cost(path_pos, goal, goal_rad) -> int:
    cost = len(path_pos) - 1
    while cost - 1 >= 0 and np.linalg.norm(path_pos[cost - 1] - goal) <= goal_rad:
        cost -= 1
    return cost


def get_travel_dist(path_pos) -> float:
    return float(sum(np.linalg.norm(path_pos[i + 1] - path_pos[i]) for i in range(len(path_pos) - 1)))
