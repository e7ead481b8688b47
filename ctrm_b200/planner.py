"""Prioritized planning over timed roadmaps on the GPU — host mirror of the reference's planner package for this step
(SURVEY.md section 8f row 2): `PrioritizedPlanning(ins, trms).solve() -> Result` (planner/prioritized_planning.py:25-101,
planner/planner.py:82-144, planner/result.py:14-40, planner/utils.py:13-54), plus the batch form the roadmap builder feeds
directly from device memory (`plan_prioritized_device`).  The search itself is `ctrm_plan_prioritized` (csrc/plan.cu);
there is no CPU fallback."""
from __future__ import annotations

import time
from dataclasses import asdict, dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from . import _lib
from .timed_roadmap import TimedNode


@dataclass(frozen=True)
class Result:
    """planner/result.py:14-40"""
    solved: bool = False
    paths: List[List[TimedNode]] = field(default_factory=list)
    name_planner: str = ""
    elapsed_planner: float = 0
    sum_of_costs: float = 0
    maximum_costs: float = 0
    sum_of_travel_dists: float = 0
    maximum_travel_dists: float = 0
    cnt_static_collide: int = 0
    cnt_continuous_collide: int = 0
    elapsed_static_collide: float = 0
    elapsed_continuous_collide: float = 0
    lowlevel_expanded: int = 0
    lowlevel_explored: int = 0

    def get_dict_wo_paths(self) -> dict:
        dic = asdict(self)
        del dic["paths"]
        return dic


def get_cost(path_pos: np.ndarray, goal: np.ndarray, goal_rad: float) -> int:
    """planner/utils.py:13-35 on an array of positions"""
    cost = len(path_pos) - 1
    while cost - 1 >= 0 and np.linalg.norm(path_pos[cost - 1] - goal) <= goal_rad:
        cost -= 1
    return cost


def get_travel_dist(path_pos: np.ndarray) -> float:
    """planner/utils.py:38-54"""
    return float(sum(np.linalg.norm(path_pos[i + 1] - path_pos[i]) for i in range(len(path_pos) - 1)))


def plan_prioritized_device(B: int, A: int, L: int, agent_off, n_layers, layer_ptr, pos, eptr, eidx, goals, speeds, rads,
                            stream: Optional[int] = None) -> dict:
    """all arguments are torch CUDA tensors in the layout of ctrm_export_csr / ctrm_upload_instances; returns CUDA tensors
    paths [A][L] i32, path_pos [A][L][2] f64, solved [B] i32, failed_agent [B] i32, counters [B][3] i64, and elapsed_ms"""
    import torch

    lib = _lib.load()
    dev = pos.device
    # vertices per agent: layer_ptr[a*L + n_layers[a]] - layer_ptr[a*L]
    base = torch.arange(A, device=dev, dtype=torch.int64) * L
    lp = layer_ptr.to(torch.int64)
    nv_agent = lp[base + n_layers.to(torch.int64)] - lp[base]
    max_nv = int(nv_agent.max().item())
    out = dict(paths=torch.empty((A, L), dtype=torch.int32, device=dev),
               path_pos=torch.zeros((A, L, 2), dtype=torch.float64, device=dev),
               solved=torch.zeros(B, dtype=torch.int32, device=dev),
               failed_agent=torch.full((B,), -1, dtype=torch.int32, device=dev),
               counters=torch.zeros((B, 3), dtype=torch.int64, device=dev))
    with torch.cuda.device(dev):  # events and the launch on the tensors' device, whatever the caller's current device is
        sobj = torch.cuda.current_stream(dev) if stream is None else torch.cuda.ExternalStream(stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sobj)  # on the stream the kernel is launched on
        _lib.check(lib.ctrm_plan_prioritized(B, A, L, agent_off.data_ptr(), n_layers.data_ptr(), layer_ptr.data_ptr(),
                                             pos.data_ptr(), eptr.data_ptr(), eidx.data_ptr(), goals.data_ptr(),
                                             speeds.data_ptr(), rads.data_ptr(), max_nv, out["paths"].data_ptr(),
                                             out["path_pos"].data_ptr(), out["solved"].data_ptr(),
                                             out["failed_agent"].data_ptr(), out["counters"].data_ptr(), sobj.cuda_stream),
                   "ctrm_plan_prioritized")
        e1.record(sobj)
        e1.synchronize()
        out["elapsed_ms"] = e0.elapsed_time(e1)
    return out


def plan_prioritized_arrays(agent_off, L: int, n_layers, layer_ptr, pos, eptr, eidx, goals, speeds, rads, device: int = 0) -> dict:
    """host (numpy) arrays in, host arrays out; same layout and results as plan_prioritized_device"""
    import torch

    dev = torch.device("cuda", device)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    B, A = len(agent_off) - 1, len(n_layers)
    r = plan_prioritized_device(B, A, int(L), up(agent_off, np.int32), up(n_layers, np.int32), up(layer_ptr, np.int32),
                                up(pos, np.float64), up(eptr, np.int64), up(eidx if len(eidx) else np.zeros(1), np.int32),
                                up(goals, np.float64), up(speeds, np.float64), up(rads, np.float64))
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in r.items()}


def _flatten_trms(trms: Sequence) -> tuple:
    """list[TimedRoadmap] of ONE instance -> (L, n_layers, layer_ptr with stride L, pos, eptr, eidx)"""
    if all(hasattr(t, "_csr") and hasattr(t, "_agent") for t in trms) and len({id(t._csr) for t in trms}) == 1:
        csr = trms[0]._csr
        ag = [t._agent for t in trms]
        if ag == list(range(ag[0], ag[0] + len(ag))):  # a contiguous run of agents of one CSR result: slice, no Python lists
            return csr.agents_arrays(ag[0], ag[-1] + 1)
    n_layers = np.array([len(t.V) for t in trms], dtype=np.int32)
    L = int(n_layers.max()) + 1
    layer_ptr = np.zeros(len(trms) * L + 1, dtype=np.int32)
    pos, eptr, eidx = [], [0], []
    for a, trm in enumerate(trms):
        for t in range(L):
            layer_ptr[a * L + t] = len(pos)
            if t < len(trm.V):
                for i, v in enumerate(trm.V[t]):
                    pos.append(np.asarray(v.pos, dtype=np.float64))
                    eidx.extend(int(j) for j in trm.E[t][i])
                    eptr.append(len(eidx))
    layer_ptr[len(trms) * L] = len(pos)
    return (L, n_layers, layer_ptr, np.array(pos, dtype=np.float64).reshape(-1, 2), np.array(eptr, dtype=np.int64),
            np.array(eidx, dtype=np.int32))


class PrioritizedPlanning:
    """drop-in for ctrm.planner.PrioritizedPlanning (prioritized_planning.py:25-101): same constructor, `solve()` returns
    a Result with the same fields; `solution` holds the paths as lists of TimedNode.  Like the reference's solve()
    (planner/planner.py:87-141) it runs validate() on a solved instance, whose pair tests count in cnt_continuous_collide;
    the batch form `plan_prioritized_device` reports the counters of _solve() alone.  `time_limit` is accepted and ignored
    (the search takes milliseconds); `agent_collision=False` is not supported."""

    def __init__(self, ins, trms, verbose: int = 0, **kwargs):
        if kwargs.get("agent_collision", True) is not True:
            raise _lib.CtrmError("PrioritizedPlanning: agent_collision=False is not supported on the GPU path")
        self.ins, self.trms, self.verbose = ins, trms, verbose
        self.device = int(kwargs.get("device", 0))
        self.solved = False
        self.solution: List[List[TimedNode]] = []
        self.elapsed = 0.0
        self.result: Optional[Result] = None
        self.cnt_static_collide = 0
        self.cnt_continuous_collide = 0
        self.lowlevel_expanded = 0
        self.lowlevel_explored = 0
        self.path_pos: Optional[np.ndarray] = None

    def get_name(self) -> str:
        return "PrioritizedPlanning"

    def validate(self) -> bool:
        """Planner.validate (planner/planner.py:150-209): starts, goals, continuity and one swept-circle test per agent pair
        and time step, in the reference's (t, i, j) order; every executed test counts in cnt_continuous_collide exactly as
        the reference's collide_dynamic_agents does (:234-238).  The pair tests run as one ctrm_collide_sphere_pairs call."""
        if not self.solved or self.path_pos is None:
            return False
        import torch

        ins, P = self.ins, self.path_pos
        N, T = P.shape[0], P.shape[1] - 1
        starts = np.asarray(ins.starts, dtype=np.float64).reshape(N, 2)
        goals = np.asarray(ins.goals, dtype=np.float64).reshape(N, 2)
        if not np.array_equal(P[:, 0], starts):
            return False
        goal_rads = np.asarray(getattr(ins, "goal_rads", np.zeros(N)), dtype=np.float64)
        if not all(np.linalg.norm(P[i, -1] - goals[i]) <= goal_rads[i] for i in range(N)):
            return False
        if T < 1:
            return True
        speeds = np.asarray(ins.max_speeds, dtype=np.float64)
        rads = np.asarray(ins.rads, dtype=np.float64)
        cont_bad = np.array([[np.linalg.norm(P[i, t] - P[i, t - 1]) > speeds[i] for i in range(N)] for t in range(1, T + 1)])
        ii, jj = np.triu_indices(N, 1)  # row-major: i ascending, then j > i ascending — the reference's loop order
        n_pairs = len(ii)
        hit = np.zeros((T, n_pairs), dtype=bool)
        if n_pairs:
            dev = torch.device("cuda", self.device)
            with torch.cuda.device(dev):
                up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
                f1, t1 = up(P[ii, :-1].transpose(1, 0, 2)), up(P[ii, 1:].transpose(1, 0, 2))  # [T][pairs][2]
                f2, t2 = up(P[jj, :-1].transpose(1, 0, 2)), up(P[jj, 1:].transpose(1, 0, 2))
                r1, r2 = up(np.tile(rads[ii], T)), up(np.tile(rads[jj], T))
                out = torch.zeros(T * n_pairs, dtype=torch.uint8, device=dev)
                _lib.check(_lib.load().ctrm_collide_sphere_pairs(f1.data_ptr(), t1.data_ptr(), r1.data_ptr(), f2.data_ptr(),
                                                                 t2.data_ptr(), r2.data_ptr(), 2, T * n_pairs, out.data_ptr(),
                                                                 torch.cuda.current_stream(dev).cuda_stream),
                           "ctrm_collide_sphere_pairs")
                hit = out.cpu().numpy().reshape(T, n_pairs).astype(bool)
        first_pair = np.searchsorted(ii, np.arange(N))  # index of agent i's first pair (i, i+1)
        for t in range(T):
            for i in range(N):
                if cont_bad[t, i]:
                    return False
                lo, hi = int(first_pair[i]), int(first_pair[i] + (N - 1 - i))
                h = np.nonzero(hit[t, lo:hi])[0]
                if len(h):
                    self.cnt_continuous_collide += int(h[0]) + 1
                    return False
                self.cnt_continuous_collide += hi - lo
        return True

    def solve(self) -> Result:
        t0 = time.time()
        ins = self.ins
        N = ins.num_agents
        L, n_layers, layer_ptr, pos, eptr, eidx = _flatten_trms(self.trms)
        if len(set(int(n) for n in n_layers)) != 1:
            raise _lib.CtrmError("PrioritizedPlanning: timed roadmaps of one instance must have equal length")
        r = plan_prioritized_arrays(np.array([0, N], dtype=np.int32), L, n_layers, layer_ptr, pos, eptr, eidx,
                                    np.asarray(ins.goals, dtype=np.float64).reshape(N, 2),
                                    np.asarray(ins.max_speeds, dtype=np.float64), np.asarray(ins.rads, dtype=np.float64), self.device)
        self.solved = bool(r["solved"][0])
        self.cnt_continuous_collide, self.lowlevel_expanded, self.lowlevel_explored = (int(x) for x in r["counters"][0])
        T = int(n_layers[0]) - 1
        kw: dict[str, Any] = {}
        if self.solved:
            self.path_pos = r["path_pos"][:, :T + 1].copy()
            if not self.validate():  # planner.py:103-105 — the reference only logs; so do we
                import logging

                logging.getLogger(__name__).error("inconsistent planner")
            self.solution = [[TimedNode(t, int(r["paths"][a, t]), self.path_pos[a, t].copy()) for t in range(T + 1)] for a in range(N)]
            goal_rads = np.asarray(getattr(ins, "goal_rads", np.zeros(N)), dtype=np.float64)
            costs = [get_cost(self.path_pos[a], np.asarray(ins.goals[a]), float(goal_rads[a])) for a in range(N)]
            dists = [get_travel_dist(self.path_pos[a]) for a in range(N)]
            kw = dict(paths=self.solution, sum_of_costs=sum(costs), maximum_costs=max(costs), sum_of_travel_dists=sum(dists),
                      maximum_travel_dists=max(dists))
        self.elapsed = time.time() - t0
        self.result = Result(solved=self.solved, name_planner=self.get_name(), elapsed_planner=self.elapsed,
                             cnt_static_collide=0, cnt_continuous_collide=self.cnt_continuous_collide,
                             elapsed_continuous_collide=r["elapsed_ms"] * 1e-3, lowlevel_expanded=self.lowlevel_expanded,
                             lowlevel_explored=self.lowlevel_explored, **kw)
        return self.result
