"""Timed roadmap container: the output type of the path (mirror of roadmap/timed_roadmap.py:16-272).

  TimedNode      (t, index, pos)                       timed_roadmap.py:16-43
  TimedRoadmap   V[t][i] -> TimedNode, E[t][i] -> indices into V[t+1]     timed_roadmap.py:45-57
  CSRRoadmaps    the GPU result of a whole batch as flat arrays (what libctrm_b200 exports)
  CompactCSRRoadmaps  the same with byte-sized counts / degrees / successor indices, as it travels to the host
  LazyTimedRoadmap   a TimedRoadmap whose V / E lists are materialised from the CSR on first access — building
                 ~42 k TimedNode objects per instance in CPython costs far more than building the roadmap on the GPU
                 (SURVEY §7.2 item 7), so consumers that can read arrays should use CSRRoadmaps directly.
Adjacency lists come out in ascending index order; the reference's insertion order (and CPython set iteration
order inside merge_samples) is not reproduced — the edge SETS are identical.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class TimedNode:
    t: int
    index: int
    pos: np.ndarray

    def __eq__(self, other) -> bool:
        return (type(other) is TimedNode and self.t == other.t and self.index == other.index
                and np.array_equal(self.pos, other.pos))

    def __getstate__(self):
        return (self.t, self.index, self.pos)

    def __setstate__(self, state):
        for k, v in zip(("t", "index", "pos"), state):
            object.__setattr__(self, k, v)


class TimedRoadmap:
    def __init__(self, loc_start: np.ndarray):
        self.V = [[TimedNode(0, 0, loc_start)]]
        self.E = [[[]]]
        self.dim = loc_start.shape[0]

    def __repr__(self) -> str:
        out = []
        for t in range(len(self.V)):
            for i, v in enumerate(self.V[t]):
                out.append(f"{i}, {v}, children={self.E[t][i]}")
            out.append("")
        return "\n".join(out) + "\n"

    def extend_layer(self, t: int) -> None:
        while t > len(self.V) - 1:
            self.V.append([])
            self.E.append([])

    def append_sample(self, loc, t, valid_edge=None, parents=None, children=None) -> None:
        assert t > 0, "use __init__ instead"
        self.extend_layer(t)
        node = TimedNode(t, len(self.V[t]), loc)
        self.V[t].append(node)
        self.E[t].append([])
        if parents is None and valid_edge is not None:
            self.add_parents(node, valid_edge)
        elif parents is not None:
            for p in parents:
                self.E[t - 1][p].append(node.index)
        if children is None and valid_edge is not None:
            self.add_children(node, valid_edge)
        elif children is not None:
            self.E[t][node.index].extend(children)

    def add_parents(self, node, valid_edge) -> None:
        if node.t < 1:
            return
        for i, u in enumerate(self.V[node.t - 1]):
            if valid_edge(u.pos, node.pos):
                self.E[node.t - 1][i].append(node.index)

    def add_children(self, node, valid_edge) -> None:
        if node.t + 1 > len(self.V) - 1:
            return
        for i, u in enumerate(self.V[node.t + 1]):
            if valid_edge(node.pos, u.pos):
                self.E[node.t][node.index].append(i)

    def append_samples(self, locs, t, valid_edge) -> None:
        for loc in locs:
            self.append_sample(loc, t, valid_edge)

    def get_parents(self, u: TimedNode):
        if u.t < 1:
            return []
        return [v.index for v in self.V[u.t - 1] if u.index in self.E[u.t - 1][v.index]]

    def getNodesPosOnestep(self, t: int) -> np.ndarray:
        if t >= len(self.V):
            return np.empty((0, self.dim))
        return np.array([u.pos for u in self.V[t]])


class CSRRoadmaps:
    """flat result of ctrm_export_csr for a batch of B instances / A agents (vertex order: agent, layer, index)"""

    def __init__(self, agent_off, L, layer_ptr, pos, eptr, eidx, n_layers, cnt_collide, timing=None):
        self.agent_off = agent_off      # (B+1,) int32
        self.L = int(L)                 # layer slots per agent (max_T + 2)
        self.layer_ptr = layer_ptr      # (A*L+1,) int32
        self.pos = pos                  # (nv, 2) float64
        self.eptr = eptr                # (nv+1,) int64
        self.eidx = eidx                # (ne,) int32, indices into the next layer, ascending
        self.n_layers = n_layers        # (A,) int32  len(trm.V)
        self.cnt_collide = cnt_collide  # (B,) int64  StaticObjects.cnt_continuous_collide
        self.timing = timing or {}

    def copy(self) -> "CSRRoadmaps":
        """detach from the builder's cached pinned staging buffers"""
        return CSRRoadmaps(self.agent_off.copy(), self.L, self.layer_ptr.copy(), self.pos.copy(), self.eptr.copy(), self.eidx.copy(),
                           self.n_layers.copy(), self.cnt_collide.copy(), dict(self.timing))

    @property
    def num_instances(self) -> int:
        return len(self.agent_off) - 1

    @property
    def n_vertices(self) -> int:
        return int(self.pos.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.eidx.shape[0])

    def instance_counts(self, b: int):
        a0, a1 = int(self.agent_off[b]), int(self.agent_off[b + 1])
        v0, v1 = int(self.layer_ptr[a0 * self.L]), int(self.layer_ptr[a1 * self.L])
        return v1 - v0, int(self.eptr[v1] - self.eptr[v0])

    def agents_arrays(self, a0: int, a1: int):
        """(L, n_layers, layer_ptr with stride L, pos, eptr, eidx) of agents [a0, a1), offsets local to the slice"""
        L = self.L
        v0, v1 = int(self.layer_ptr[a0 * L]), int(self.layer_ptr[a1 * L])
        e0, e1 = int(self.eptr[v0]), int(self.eptr[v1])
        return (L, self.n_layers[a0:a1], self.layer_ptr[a0 * L:a1 * L + 1] - v0, self.pos[v0:v1], self.eptr[v0:v1 + 1] - e0,
                self.eidx[e0:e1])

    def instance_arrays(self, b: int):
        return self.agents_arrays(int(self.agent_off[b]), int(self.agent_off[b + 1]))

    def layer(self, agent: int, t: int):
        """(positions (n,2), list of adjacency arrays) of V[t] / E[t] for flat agent index `agent`"""
        v0, v1 = int(self.layer_ptr[agent * self.L + t]), int(self.layer_ptr[agent * self.L + t + 1])
        return self.pos[v0:v1], [self.eidx[self.eptr[v]:self.eptr[v + 1]] for v in range(v0, v1)]

    def timed_roadmaps(self, b: int, lazy: bool = True):
        a0, a1 = int(self.agent_off[b]), int(self.agent_off[b + 1])
        cls = LazyTimedRoadmap if lazy else _materialise
        return [cls(self, a) for a in range(a0, a1)]


class CompactCSRRoadmaps(CSRRoadmaps):
    """the same result as it arrives from ctrm_export_csr_compact: vertices per layer, out-degree per vertex and successor
    index as single bytes, positions fp64, plus the first vertex / edge of every agent.  Per-instance and per-agent views are
    expanded from the bytes on demand (a prefix sum over ~1,300 vertices); `layer_ptr`, `eptr` and `eidx` of the WHOLE batch in
    the wide layout of CSRRoadmaps are computed once on first access for callers that want them."""

    def __init__(self, agent_off, L, layer_cnt, pos, deg, eidx8, agent_v0, agent_e0, n_layers, cnt_collide, timing=None):
        self.agent_off = agent_off
        self.L = int(L)
        self.layer_cnt = layer_cnt      # (A*L,) uint8
        self.pos = pos                  # (nv, 2) float64
        self.deg = deg                  # (nv,) uint8
        self.eidx8 = eidx8              # (ne,) uint8
        self.agent_v0 = agent_v0        # (A+1,) int64
        self.agent_e0 = agent_e0        # (A+1,) int64
        self.n_layers = n_layers
        self.cnt_collide = cnt_collide
        self.timing = timing or {}
        self._wide = {}

    def copy(self) -> "CompactCSRRoadmaps":
        return CompactCSRRoadmaps(self.agent_off.copy(), self.L, self.layer_cnt.copy(), self.pos.copy(), self.deg.copy(),
                                  self.eidx8.copy(), self.agent_v0.copy(), self.agent_e0.copy(), self.n_layers.copy(),
                                  self.cnt_collide.copy(), dict(self.timing))

    @property
    def nbytes(self) -> int:
        return int(self.layer_cnt.nbytes + self.pos.nbytes + self.deg.nbytes + self.eidx8.nbytes + self.agent_v0.nbytes +
                   self.agent_e0.nbytes + self.n_layers.nbytes + self.cnt_collide.nbytes)

    @property
    def layer_ptr(self):
        if "layer_ptr" not in self._wide:
            lp = np.zeros(len(self.layer_cnt) + 1, dtype=np.int32)
            np.cumsum(self.layer_cnt, dtype=np.int32, out=lp[1:])
            self._wide["layer_ptr"] = lp
        return self._wide["layer_ptr"]

    @property
    def eptr(self):
        if "eptr" not in self._wide:
            ep = np.zeros(len(self.deg) + 1, dtype=np.int64)
            np.cumsum(self.deg, dtype=np.int64, out=ep[1:])
            self._wide["eptr"] = ep
        return self._wide["eptr"]

    @property
    def eidx(self):
        if "eidx" not in self._wide:
            self._wide["eidx"] = self.eidx8.astype(np.int32)
        return self._wide["eidx"]

    @property
    def n_edges(self) -> int:
        return int(self.eidx8.shape[0])

    def instance_counts(self, b: int):
        a0, a1 = int(self.agent_off[b]), int(self.agent_off[b + 1])
        return int(self.agent_v0[a1] - self.agent_v0[a0]), int(self.agent_e0[a1] - self.agent_e0[a0])

    def agents_arrays(self, a0: int, a1: int):
        L = self.L
        v0, v1 = int(self.agent_v0[a0]), int(self.agent_v0[a1])
        e0, e1 = int(self.agent_e0[a0]), int(self.agent_e0[a1])
        lp = np.zeros((a1 - a0) * L + 1, dtype=np.int32)
        np.cumsum(self.layer_cnt[a0 * L:a1 * L], dtype=np.int32, out=lp[1:])
        ep = np.zeros(v1 - v0 + 1, dtype=np.int64)
        np.cumsum(self.deg[v0:v1], dtype=np.int64, out=ep[1:])
        return L, self.n_layers[a0:a1], lp, self.pos[v0:v1], ep, self.eidx8[e0:e1].astype(np.int32)

    def layer(self, agent: int, t: int):
        _, _, lp, pos, ep, ei = self.agents_arrays(agent, agent + 1)
        v0, v1 = int(lp[t]), int(lp[t + 1])
        return pos[v0:v1], [ei[ep[v]:ep[v + 1]] for v in range(v0, v1)]


def _build_lists(csr: CSRRoadmaps, agent: int):
    V, E = [], []
    _, _, lp, pos, ep, ei = csr.agents_arrays(agent, agent + 1)
    for t in range(int(csr.n_layers[agent])):
        v0, v1 = int(lp[t]), int(lp[t + 1])
        V.append([TimedNode(t, i, pos[v0 + i].copy()) for i in range(v1 - v0)])
        E.append([ei[ep[v]:ep[v + 1]].tolist() for v in range(v0, v1)])
    return V, E


def _materialise(csr: CSRRoadmaps, agent: int) -> TimedRoadmap:
    trm = TimedRoadmap.__new__(TimedRoadmap)
    trm.V, trm.E = _build_lists(csr, agent)
    trm.dim = 2
    return trm


class LazyTimedRoadmap(TimedRoadmap):
    """TimedRoadmap-compatible view over a CSRRoadmaps result; V and E are built on first access"""

    def __init__(self, csr: CSRRoadmaps, agent: int):
        self._csr = csr
        self._agent = agent
        self.dim = 2

    def __getattr__(self, name):
        if name in ("V", "E"):
            V, E = _build_lists(self._csr, self._agent)
            self.__dict__["V"] = V
            self.__dict__["E"] = E
            return self.__dict__[name]
        raise AttributeError(name)

    def __len__(self):
        return int(self._csr.n_layers[self._agent])
