"""Baseline samplers on the GPU collision / edge kernels — host mirror of the reference's `ctrm.roadmap` package for the
samplers that are pure distance + swept-collision + prefix-scan problems:

  get_random_samples / get_timed_roadmaps_random / _random_common / _fully_random   roadmap/random_sampler.py:22-124
  get_grid / get_timed_roadmaps_grid / _grid_common / _grid_common_2d_fast           roadmap/grid_sampler.py:19-136
  valid_move / get_fixed_structured_roadmaps / get_common_roadmaps                   roadmap/utils.py:18-134
  TimedRoadmap.append_fixed_strucutre                                                roadmap/timed_roadmap.py:169-210

Same names, arguments, RNG consumption (numpy's global stream, one draw of `dim` doubles per candidate, in the reference's
order) and side effects (`ins.objs.cnt_static_collide` / `cnt_continuous_collide`).  The device work is three C-ABI calls:
`ctrm_collide_points` (static point-vs-circle test for all candidates at once), `ctrm_pair_adjacency` (all-pairs
valid_move as bit matrices, every agent's sample set in one launch) and `ctrm_adjacency_csr` (popcount + prefix scan +
scatter).  Adjacency lists come out in the reference's own order: `[i]` first, then the neighbours ascending.  No CPU fallback.
"""
from __future__ import annotations

from itertools import product
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .timed_roadmap import TimedNode, TimedRoadmap


# ---------------------------------------------------------------------------------------------- device helpers
def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.CtrmError("ctrm_b200.roadmap needs a CUDA device (no CPU fallback)")
    return torch


def _dev(a, dtype):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def _obs_dev(ins):
    xyr = ins.objs._xyr() if hasattr(ins.objs, "_xyr") else np.array(
        [[o.pos[0], o.pos[1], o.rad] for o in ins.obs], dtype=np.float64).reshape(-1, 3)
    return _dev(xyr if len(xyr) else np.zeros((1, 3)), np.float64), _dev(np.array([0, len(xyr)]), np.int32)


def collide_points(ins, pts: np.ndarray, rad: float) -> np.ndarray:
    """StaticObjects.collide_sphere (static_objects.py:44-57) for every row of pts at once; bumps cnt_static_collide by
    len(pts) like the reference's one call per candidate"""
    torch = _torch()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    n = len(pts)
    if n == 0:
        return np.zeros(0, dtype=bool)
    d_obs, d_off = _obs_dev(ins)
    d_pts, d_rad = _dev(pts, np.float64), _dev(np.full(n, float(rad)), np.float64)
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().ctrm_collide_points(d_pts.data_ptr(), d_rad.data_ptr(), None, n, d_obs.data_ptr(), d_off.data_ptr(),
                                               out.data_ptr(), _stream()), "ctrm_collide_points")
    ins.objs.cnt_static_collide += n
    return out.cpu().numpy().astype(bool)


def pair_adjacency(ins, rows: Sequence[np.ndarray], cols: Sequence[np.ndarray], speeds, rads, symmetric: bool, self_first: bool):
    """all-pairs valid_move for several (rows_s, cols_s) blocks at once.  Returns (eptr (R+1,) int64, eidx int32, counts
    (n_sets,) int64): the adjacency lists of all rows, set after set, and the number of swept tests per set."""
    torch = _torch()
    lib = _lib.load()
    n_sets = len(rows)
    row_off = np.zeros(n_sets + 1, dtype=np.int32)
    col_off = np.zeros(n_sets + 1, dtype=np.int32)
    row_off[1:] = np.cumsum([len(r) for r in rows])
    col_off[1:] = np.cumsum([len(c) for c in cols])
    R = int(row_off[-1])
    if n_sets == 0 or R == 0:
        return np.zeros(R + 1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(n_sets, dtype=np.int64)
    speeds = np.ascontiguousarray(speeds, dtype=np.float64)
    rads = np.ascontiguousarray(rads, dtype=np.float64)
    speeds2 = np.array([float(s) ** 2 for s in speeds], dtype=np.float64)  # max_speed ** 2 is the caller's own pow
    cat = lambda xs: np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1, 2) for x in xs] + [np.zeros((1, 2))])  # noqa: E731
    d_rows = _dev(cat(rows), np.float64)
    d_cols = d_rows if symmetric else _dev(cat(cols), np.float64)
    d_obs, d_off = _obs_dev(ins)
    words = int(lib.ctrm_pair_adjacency_words(_lib.ptr(row_off), _lib.ptr(col_off), n_sets))
    bits = torch.empty(max(words, 1), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(n_sets, dtype=torch.int64, device="cuda")
    st = _stream()
    _lib.check(lib.ctrm_pair_adjacency(d_rows.data_ptr(), _lib.ptr(row_off), d_cols.data_ptr(), _lib.ptr(col_off), n_sets, None,
                                       _lib.ptr(speeds), _lib.ptr(speeds2), _lib.ptr(rads), d_obs.data_ptr(), d_off.data_ptr(),
                                       1 if symmetric else 0, bits.data_ptr(), cnt.data_ptr(), st), "ctrm_pair_adjacency")
    eptr = torch.empty(R + 1, dtype=torch.int64, device="cuda")
    _lib.check(lib.ctrm_adjacency_csr(bits.data_ptr(), _lib.ptr(row_off), _lib.ptr(col_off), n_sets, 1 if self_first else 0,
                                      eptr.data_ptr(), None, st), "ctrm_adjacency_csr")
    ne = int(eptr[R].item())
    eidx = torch.empty(max(ne, 1), dtype=torch.int32, device="cuda")
    _lib.check(lib.ctrm_adjacency_csr(bits.data_ptr(), _lib.ptr(row_off), _lib.ptr(col_off), n_sets, 1 if self_first else 0,
                                      eptr.data_ptr(), eidx.data_ptr(), st), "ctrm_adjacency_csr")
    return eptr.cpu().numpy(), eidx[:ne].cpu().numpy(), cnt.cpu().numpy()


def _lists(eptr, eidx, r0, r1):
    return [eidx[eptr[r]:eptr[r + 1]].tolist() for r in range(r0, r1)]


def valid_move(pos1, pos2, max_speed: float, rad: float, objs) -> bool:
    """roadmap/utils.py:18-42 for one pair (the batch paths above are what the samplers use)"""
    p = np.abs(np.asarray(pos1, dtype=np.float64) - np.asarray(pos2, dtype=np.float64))
    if any(p > max_speed):
        return False
    if sum(p ** 2) > max_speed ** 2:
        return False
    return not objs.collide_continuous_sphere(pos1, pos2, rad)


# ---------------------------------------------------------------------------------------------- containers
class FixedStructureRoadmap(TimedRoadmap):
    """TimedRoadmap after append_fixed_strucutre (timed_roadmap.py:169-210): V[0] = [start], V[1..T] = the same `locs`,
    E[0][0] = the locations reachable from the start, E[1..T-1] = ONE shared adjacency (the reference assigns the same
    list object to every layer), E[T] = empty lists.  `structure` exposes the arrays; V / E are materialised on first access."""

    def __init__(self, start, locs: np.ndarray, T: int, start_edges: List[int], adjs: List[List[int]]):
        self.dim = 2
        self._start = np.asarray(start, dtype=np.float64)
        self._locs = np.asarray(locs, dtype=np.float64).reshape(-1, 2)
        self._T = int(T)
        self._start_edges = start_edges
        self._adjs = adjs

    @property
    def structure(self) -> dict:
        return dict(start=self._start, locs=self._locs, T=self._T, start_edges=self._start_edges, adjs=self._adjs)

    def __len__(self):
        return self._T + 1

    def __getattr__(self, name):
        if name in ("V", "E"):
            n, T = len(self._locs), self._T
            V = [[TimedNode(0, 0, self._start)]] + [[TimedNode(t, i, self._locs[i]) for i in range(n)] for t in range(1, T + 1)]
            E = [[list(self._start_edges)]] + [self._adjs for _ in range(1, T)] + ([[[] for _ in range(n)]] if T >= 1 else [])
            if T == 1:
                E = [[list(self._start_edges)], [[] for _ in range(n)]]
            self.__dict__["V"], self.__dict__["E"] = V, E
            return self.__dict__[name]
        raise AttributeError(name)


# ---------------------------------------------------------------------------------------------- samplers
def _draw_until(num: int, rad: float, ins, dim: int) -> np.ndarray:
    """`while len(locs) < num: pos = np.random.random(dim) * (1 - 2 * rad) + rad; keep unless collide_sphere`
    (random_sampler.py:39-43) with the candidates tested in batches: the accepted positions, the number of candidates
    consumed and numpy's global stream afterwards are exactly the sequential loop's"""
    state = np.random.get_state()
    kept, consumed, cands_seen = [], 0, 0
    need = num
    while need > 0:
        m = int(need * 1.5) + 32
        cand = np.random.random((m, dim)) * (1 - 2 * rad) + rad  # row k = the k-th sequential np.random.random(dim)
        free = ~_collide_points_nocount(ins, cand, rad)
        idx = np.nonzero(free)[0]
        if len(idx) >= need:
            cut = int(idx[need - 1]) + 1
            kept.append(cand[idx[:need]])
            consumed = cands_seen + cut
            need = 0
        else:
            kept.append(cand[idx])
            need -= len(idx)
            cands_seen += m
    np.random.set_state(state)
    if consumed:
        np.random.random((consumed, dim))  # leave the stream where the reference's loop leaves it
    ins.objs.cnt_static_collide += consumed
    return np.concatenate(kept) if kept else np.zeros((0, dim))


def _collide_points_nocount(ins, pts, rad):
    before = ins.objs.cnt_static_collide
    v = collide_points(ins, pts, rad)
    ins.objs.cnt_static_collide = before
    return v


def get_random_samples(num: int, rad: float, ins, wo_invalid_samples: bool = True) -> List[np.ndarray]:
    """random_sampler.py:22-51"""
    dim = int(getattr(ins, "dim", 2))
    if dim != 2:
        raise _lib.CtrmError("baseline samplers: only dim = 2 is supported (README.md:129)")
    if wo_invalid_samples:
        return list(_draw_until(num, rad, ins, dim))
    cand = np.random.random((num, dim)) * (1 - 2 * rad) + rad
    return list(cand[~collide_points(ins, cand, rad)])


def get_grid(size: int, rad: float, ins) -> List[np.ndarray]:
    """grid_sampler.py:19-36"""
    pts = np.array([s for s in product(np.linspace(rad, 1 - rad, size), repeat=2)], dtype=np.float64).reshape(-1, 2)
    return list(pts[~collide_points(ins, pts, rad)])


def get_fixed_structured_roadmaps(ins, T: int, arr_locs: Sequence[Sequence[np.ndarray]], verbose: int = 0) -> List[TimedRoadmap]:
    """roadmap/utils.py:45-84 + TimedRoadmap.append_fixed_strucutre (timed_roadmap.py:169-210), all agents in two launches"""
    N = ins.num_agents
    locs = [np.asarray(l, dtype=np.float64).reshape(-1, 2) for l in arr_locs]
    speeds, rads = list(ins.max_speeds), list(ins.rads)
    starts = [np.asarray(ins.starts[i], dtype=np.float64).reshape(1, 2) for i in range(N)]
    sp, se, scnt = pair_adjacency(ins, starts, locs, speeds, rads, symmetric=False, self_first=False)   # add_parents of layer 1
    ap, ae, acnt = pair_adjacency(ins, locs, locs, speeds, rads, symmetric=True, self_first=True)       # adjs
    ins.objs.cnt_continuous_collide += int(scnt.sum() + acnt.sum())
    trms, r0 = [], 0
    for i in range(N):
        n = len(locs[i])
        trms.append(FixedStructureRoadmap(ins.starts[i], locs[i], T, se[sp[i]:sp[i + 1]].tolist(), _lists(ap, ae, r0, r0 + n)))
        r0 += n
    return trms


def get_common_roadmaps(ins, T: int, locs_common: Sequence[np.ndarray], adjs_common: Optional[List[List[int]]] = None) -> List[TimedRoadmap]:
    """roadmap/utils.py:87-134"""
    assert all(rad == ins.rads[0] for rad in ins.rads)
    assert all(speed == ins.max_speeds[0] for speed in ins.max_speeds)
    N = ins.num_agents
    speed, rad = ins.max_speeds[0], ins.rads[0]
    lc = np.asarray(locs_common, dtype=np.float64).reshape(-1, 2)
    n = len(lc)
    if adjs_common is None:
        ap, ae, cnt = pair_adjacency(ins, [lc], [lc], [speed], [rad], symmetric=True, self_first=True)
        ins.objs.cnt_continuous_collide += int(cnt.sum())
        adjs_common = _lists(ap, ae, 0, n)
    goals = [np.asarray(ins.goals[i], dtype=np.float64).reshape(1, 2) for i in range(N)]
    starts = [np.asarray(ins.starts[i], dtype=np.float64).reshape(1, 2) for i in range(N)]
    # valid_edge(loc_j, goal_i) for every j: rows = the common locations, one column = the goal
    gp, ge, gcnt = pair_adjacency(ins, [lc] * N, goals, [speed] * N, [rad] * N, symmetric=False, self_first=False)
    # add_parents of layer 1: valid_edge(start_i, loc) over locs_common + [goal_i]
    with_goal = [np.concatenate([lc, goals[i]]) for i in range(N)]
    sp, se, scnt = pair_adjacency(ins, starts, with_goal, [speed] * N, [rad] * N, symmetric=False, self_first=False)
    ins.objs.cnt_continuous_collide += int(gcnt.sum() + scnt.sum())
    trms = []
    for i in range(N):
        adjs = [list(a) for a in adjs_common] + [[]]
        for j in range(n):
            if gp[i * n + j + 1] > gp[i * n + j]:
                adjs[j].append(n)
                adjs[-1].append(j)
        trms.append(FixedStructureRoadmap(ins.starts[i], with_goal[i], T, se[sp[i]:sp[i + 1]].tolist(), adjs))
    return trms


def get_timed_roadmaps_random(ins, T: int, num: int, wo_invalid_samples: bool = True, verbose: int = 0) -> List[TimedRoadmap]:
    """random_sampler.py:54-78"""
    arr_locs = [get_random_samples(num, ins.rads[i], ins, wo_invalid_samples) + [np.asarray(ins.goals[i], dtype=np.float64)]
                for i in range(ins.num_agents)]
    return get_fixed_structured_roadmaps(ins, T, arr_locs, verbose)


def get_timed_roadmaps_random_common(ins, T: int, num: int, wo_invalid_samples: bool = True) -> List[TimedRoadmap]:
    """random_sampler.py:126-143"""
    return get_common_roadmaps(ins, T, get_random_samples(num, ins.rads[0], ins, wo_invalid_samples))


def get_timed_roadmaps_grid(ins, T: int, size: int, verbose: int = 0) -> List[TimedRoadmap]:
    """grid_sampler.py:39-55"""
    arr_locs = [get_grid(size, ins.rads[i], ins) + [np.asarray(ins.goals[i], dtype=np.float64)] for i in range(ins.num_agents)]
    return get_fixed_structured_roadmaps(ins, T, arr_locs, verbose)


def get_timed_roadmaps_grid_common(ins, T: int, size: int) -> List[TimedRoadmap]:
    """grid_sampler.py:58-76: in 2-D the reference takes its windowed variant (get_timed_roadmaps_grid_common_2d_fast)"""
    return get_timed_roadmaps_grid_common_2d_fast(ins, T, size)


def get_timed_roadmaps_grid_common_2d_fast(ins, T: int, size: int) -> List[TimedRoadmap]:
    """grid_sampler.py:79-136: only cell pairs within `reachable_len = int(speed * size)` cells of each other are tested (a
    pair outside that window gets no edge even where valid_move would hold); the pre-checks of the window pairs run
    vectorised on the host in the reference's fp64 order, the swept tests in one ctrm_collide_segments call"""
    torch = _torch()
    rad, speed = ins.rads[0], ins.max_speeds[0]
    reach = int(speed / (1 / size))
    pts = np.array([s for s in product(np.linspace(rad, 1 - rad, size), repeat=2)], dtype=np.float64).reshape(-1, 2)
    free = ~collide_points(ins, pts, rad)
    indexes = np.full(size * size, -1, dtype=np.int64)
    indexes[free] = np.arange(int(free.sum()))
    indexes = indexes.reshape(size, size)
    lc = pts[free]
    n = len(lc)
    # window pairs in the reference's loop order: (i, j) outer, (_i in i..i+reach, _j in j-reach..j+reach) inner, _idx > idx
    ii, jj = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    pa, pb = [], []
    for di in range(0, reach + 1):
        for dj in range(-reach, reach + 1):
            if di == 0 and dj == 0:
                continue
            i2, j2 = ii + di, jj + dj
            ok = (i2 <= size - 1) & (j2 >= 0) & (j2 <= size - 1)
            a = indexes[ii[ok], jj[ok]]
            b = indexes[i2[ok], j2[ok]]
            keep = (a >= 0) & (b >= a)  # `_idx < idx: continue` also skips the non-samples (-1)
            keep &= b != a
            pa.append(a[keep])
            pb.append(b[keep])
    pa = np.concatenate(pa) if pa else np.zeros(0, dtype=np.int64)
    pb = np.concatenate(pb) if pb else np.zeros(0, dtype=np.int64)
    p = np.abs(lc[pa] - lc[pb])
    pre = ~(p > speed).any(axis=1) & ~((0 + p[:, 0] ** 2) + p[:, 1] ** 2 > speed ** 2)
    pa, pb = pa[pre], pb[pre]
    valid = np.zeros(len(pa), dtype=bool)
    if len(pa):
        d_obs, d_off = _obs_dev(ins)
        f, t = _dev(lc[pa], np.float64), _dev(lc[pb], np.float64)
        r = _dev(np.full(len(pa), float(rad)), np.float64)
        out = torch.zeros(len(pa), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().ctrm_collide_segments(f.data_ptr(), t.data_ptr(), r.data_ptr(), None, len(pa), d_obs.data_ptr(),
                                                     d_off.data_ptr(), out.data_ptr(), None, None, _stream()), "ctrm_collide_segments")
        valid = ~out.cpu().numpy().astype(bool)
        ins.objs.cnt_continuous_collide += len(pa)
    nbr = [[] for _ in range(n)]
    for a, b in zip(pa[valid].tolist(), pb[valid].tolist()):
        nbr[a].append(b)
        nbr[b].append(a)
    adjs_common = [[i] + sorted(nbr[i]) for i in range(n)]
    return get_common_roadmaps(ins, T, list(lc), adjs_common)


def get_timed_roadmaps_fully_random(ins, T: int, num: int, wo_invalid_samples: bool = True) -> List[TimedRoadmap]:
    """random_sampler.py:81-123: fresh samples for every (t, agent), each layer wired to the previous one by add_parents
    (timed_roadmap.py:116-131) — N * T rectangular blocks in one launch"""
    N = ins.num_agents
    layers = [[np.asarray(ins.starts[i], dtype=np.float64).reshape(1, 2)] for i in range(N)]
    for t in range(1, T + 1):
        for i in range(N):
            locs = get_random_samples(num, ins.rads[i], ins, wo_invalid_samples) + [np.asarray(ins.goals[i], dtype=np.float64)]
            layers[i].append(np.asarray(locs, dtype=np.float64).reshape(-1, 2))
    rows = [layers[i][t - 1] for i in range(N) for t in range(1, T + 1)]
    cols = [layers[i][t] for i in range(N) for t in range(1, T + 1)]
    speeds = [ins.max_speeds[i] for i in range(N) for _ in range(T)]
    rads = [ins.rads[i] for i in range(N) for _ in range(T)]
    ep, ei, cnt = pair_adjacency(ins, rows, cols, speeds, rads, symmetric=False, self_first=False)
    ins.objs.cnt_continuous_collide += int(cnt.sum())
    trms, r0 = [], 0
    for i in range(N):
        trm = TimedRoadmap(np.asarray(ins.starts[i], dtype=np.float64))
        for t in range(1, T + 1):
            n_prev = len(layers[i][t - 1])
            trm.E[t - 1] = _lists(ep, ei, r0, r0 + n_prev)
            r0 += n_prev
            trm.V.append([TimedNode(t, k, layers[i][t][k]) for k in range(len(layers[i][t]))])
            trm.E.append([[] for _ in range(len(layers[i][t]))])
        trms.append(trm)
    return trms
