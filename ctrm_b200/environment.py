"""Problem definition mirror of ctrm.environment (the hot-path half).

  Instance        environment/instance.py:17-46   frozen dataclass; pickles without `objs`, rebuilt on unpickle
  ObstacleSphere  environment/obstacle.py:40-81   circle obstacle + draw_2d
  StaticObjects   environment/static_objects.py:29-138, :173-187  swept-circle test against all obstacles + draw_2d,
                  with the cnt_continuous_collide / time_continuous_collide counters eval.py reads (scripts/eval.py:83-86)
  continuous_collide_spheres  environment/fcl_utils.py:185-212

Collision and raster work runs on the GPU through libctrm_b200 (no FCL, no CPU fallback).  Pickles written by the
reference (`ctrm.environment.instance.Instance`) load through `load_instance`, which maps the reference's module
path onto these classes.
"""
from __future__ import annotations

import io
import pickle
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib


def continuous_collide_spheres(from1, to1, rad1, from2, to2, rad2) -> bool:
    """two linearly moving spheres, closest approach (fcl_utils.py:185-212 -> collision_check.cpp:16-58)"""
    import ctypes as C

    lib = _lib.load()
    f1 = np.ascontiguousarray(from1, dtype=np.float64)
    t1 = np.ascontiguousarray(to1, dtype=np.float64)
    f2 = np.ascontiguousarray(from2, dtype=np.float64)
    t2 = np.ascontiguousarray(to2, dtype=np.float64)
    out = C.c_int(0)
    _lib.check(lib.ctrm_host_continuous_collide_spheres(_lib.ptr(f1), _lib.ptr(t1), float(rad1), _lib.ptr(f2), _lib.ptr(t2),
                                                        float(rad2), int(f1.shape[0]), C.byref(out)),
               "continuousCollideSpheres")
    return bool(out.value)


@dataclass(frozen=True)
class ObstacleSphere:
    pos: np.ndarray
    rad: float

    def __post_init__(self):
        object.__setattr__(self, "circumradius", self.rad)

    def draw_2d(self, map_size: int) -> np.ndarray:
        return StaticObjects([self]).draw_2d(map_size)


class StaticObjects:
    def __init__(self, obs):
        self.obs = list(obs)
        self.cnt_static_collide = 0
        self.time_static_collide = 0.0
        self.cnt_continuous_collide = 0
        self.time_continuous_collide = 0.0

    def _xyr(self) -> np.ndarray:
        if not self.obs:
            return np.zeros((0, 3), dtype=np.float64)
        return np.array([[o.pos[0], o.pos[1], o.rad] for o in self.obs], dtype=np.float64)

    def collide_continuous_sphere(self, pos1, pos2, rad) -> bool:
        """static_objects.py:59-75, :108-138: any over obstacles of the swept-circle test; bumps the counters"""
        import torch

        t0 = time.time()
        lib = _lib.load()
        res = False
        if self.obs:
            dev = torch.device("cuda")
            seg = torch.tensor([list(pos1[:2]), list(pos2[:2])], dtype=torch.float64, device=dev)
            r = torch.tensor([float(rad)], dtype=torch.float64, device=dev)
            obs = torch.from_numpy(self._xyr()).to(dev)
            off = torch.tensor([0, len(self.obs)], dtype=torch.int32, device=dev)
            verdict = torch.zeros(1, dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.ctrm_collide_segments(seg[0].data_ptr(), seg[1].data_ptr(), r.data_ptr(), None, 1,
                                                 obs.data_ptr(), off.data_ptr(), verdict.data_ptr(), None, None, st),
                       "ctrm_collide_segments")
            res = bool(verdict.item())
        self.cnt_continuous_collide += 1
        self.time_continuous_collide += time.time() - t0
        return res

    def draw_2d(self, map_size: int) -> np.ndarray:
        """(M, M) float32 occupancy, first axis x (static_objects.py:173-187)"""
        import torch

        lib = _lib.load()
        dev = torch.device("cuda")
        obs = torch.from_numpy(self._xyr()).to(dev)
        off = torch.tensor([0, len(self.obs)], dtype=torch.int32, device=dev)
        occ = torch.zeros(map_size * map_size + 16, dtype=torch.uint8, device=dev)
        _lib.check(lib.ctrm_raster_occupancy(obs.data_ptr(), off.data_ptr(), 1, int(map_size), occ.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "ctrm_raster_occupancy")
        return occ[:map_size * map_size].reshape(map_size, map_size).cpu().numpy().astype(np.float32)


@dataclass(frozen=True)
class Instance:
    num_agents: int
    starts: list
    goals: list
    max_speeds: list
    rads: list
    goal_rads: list
    obs: list = field(default_factory=list)
    dim: int = 2
    objs: StaticObjects = field(init=False)

    def __post_init__(self):
        object.__setattr__(self, "objs", StaticObjects(self.obs))

    def __getstate__(self):
        state = self.__dict__.copy()
        del state["objs"]
        state.pop("_arrays", None)
        state.pop("_speeds2", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__post_init__()

    # ---- plain-array views used by the batch path ----
    def arrays(self):
        """(starts (N,2), goals (N,2), speeds (N,), rads (N,), obs_xyr (O,3)) float64; computed once per instance (frozen)"""
        c = self.__dict__.get("_arrays")
        if c is not None:
            return c
        try:  # fast path: homogeneous lists of (dim,) arrays / floats
            s = np.array(self.starts, dtype=np.float64).reshape(self.num_agents, -1)[:, :2]
            g = np.array(self.goals, dtype=np.float64).reshape(self.num_agents, -1)[:, :2]
            sp = np.array(self.max_speeds, dtype=np.float64).reshape(-1)
            r = np.array(self.rads, dtype=np.float64).reshape(-1)
            o = np.empty((len(self.obs), 3), dtype=np.float64)
            for k, ob in enumerate(self.obs):
                o[k, 0] = ob.pos[0]
                o[k, 1] = ob.pos[1]
                o[k, 2] = ob.rad
            s, g = np.ascontiguousarray(s), np.ascontiguousarray(g)
        except Exception:
            s = np.array([np.asarray(p, dtype=np.float64)[:2] for p in self.starts], dtype=np.float64).reshape(-1, 2)
            g = np.array([np.asarray(p, dtype=np.float64)[:2] for p in self.goals], dtype=np.float64).reshape(-1, 2)
            sp = np.array([float(v) for v in self.max_speeds], dtype=np.float64)
            r = np.array([float(v) for v in self.rads], dtype=np.float64)
            o = np.array([[float(ob.pos[0]), float(ob.pos[1]), float(ob.rad)] for ob in self.obs], dtype=np.float64).reshape(-1, 3)
        c = (s, g, sp, r, o)
        object.__setattr__(self, "_arrays", c)
        return c

    def speeds2(self) -> np.ndarray:
        """max_speed ** 2 per agent with CPython's float pow, exactly as roadmap/utils.py:40 evaluates it (cached)"""
        c = self.__dict__.get("_speeds2")
        if c is None:
            c = np.array([float(v) ** 2 for v in self.max_speeds], dtype=np.float64)
            object.__setattr__(self, "_speeds2", c)
        return c

    @staticmethod
    def from_arrays(starts, goals, max_speeds, rads, obs_pos, obs_rad, goal_rads=None) -> "Instance":
        n = len(starts)
        return Instance(num_agents=n, starts=[np.array(p, dtype=np.float64) for p in starts],
                        goals=[np.array(p, dtype=np.float64) for p in goals],
                        max_speeds=[float(v) for v in max_speeds], rads=[float(v) for v in rads],
                        goal_rads=[float(v) for v in (goal_rads if goal_rads is not None else rads)],
                        obs=[ObstacleSphere(np.array(p, dtype=np.float64), float(r)) for p, r in zip(obs_pos, obs_rad)])


class _InstanceUnpickler(pickle.Unpickler):
    """maps the reference's class paths (ctrm.environment.*) onto the mirrors above"""

    _MAP = {"Instance": Instance, "ObstacleSphere": ObstacleSphere, "StaticObjects": StaticObjects}

    def find_class(self, module, name):
        if module.split(".")[0] == "ctrm" and name in self._MAP:
            return self._MAP[name]
        return super().find_class(module, name)


def load_instance(path: str) -> Instance:
    """unpickle an instance written by the reference (e.g. notebooks/CTRM_demo_ins.pkl) or by this package"""
    with open(path, "rb") as f:
        return _InstanceUnpickler(io.BytesIO(f.read())).load()


# ---- synthetic aamas22-shape instances (config 2/3 of BASELINE.json) ----
def generate_ins_2d_with_obs_hetero_nonfix_agents(num_agents_min: int, num_agents_max: int, max_speeds_cands="0.03125",
                                                  rads_cands="0.015625", obs_num: int = 10, obs_size_lower_bound: float = 0.05,
                                                  obs_size_upper_bound: float = 0.08, rng: np.random.RandomState = None) -> Instance:
    """Synthetic instance with the shape of the aamas22 benchmarks (environment/instance.py:52-180;
    scripts/conf/data_gen/instance/hetero.yaml): N ~ randint(min, max), obstacle centres ~ U[0,1]^2 with radius
    ~ U[lb/2, ub/2], starts/goals rejection-sampled collision-free and pairwise separated.  The draw order is this
    package's own (instances are synthetic, not the reference's exact benchmark files)."""
    rng = rng or np.random.RandomState(0)
    speeds_c = [float(x) for x in str(max_speeds_cands).split(",")]
    rads_c = [float(x) for x in str(rads_cands).split(",")]
    n = int(rng.randint(num_agents_min, num_agents_max))
    obs_pos = rng.rand(obs_num, 2)
    obs_rad = rng.rand(obs_num) * (obs_size_upper_bound - obs_size_lower_bound) / 2 + obs_size_lower_bound / 2
    speeds = [speeds_c[rng.randint(len(speeds_c))] for _ in range(n)]
    rads = [rads_c[rng.randint(len(rads_c))] for _ in range(n)]

    def sample_points(existing_sets):
        pts = []
        for i in range(n):
            for _ in range(100000):
                p = rng.rand(2) * (1 - 2 * rads[i]) + rads[i]
                if obs_num and np.any(np.linalg.norm(obs_pos - p, axis=1) <= obs_rad + rads[i]):
                    continue
                if any(np.linalg.norm(p - q) <= rads[i] + rads[j] for j, q in enumerate(pts)):
                    continue
                pts.append(p)
                break
            else:
                raise RuntimeError("could not place agents")
        return pts

    starts = sample_points(None)
    goals = sample_points(None)
    return Instance.from_arrays(starts, goals, speeds, rads, obs_pos, obs_rad)
