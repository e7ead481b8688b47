"""GPU replacement of cpp_helper/cost_to_go_wrapper (cost_to_go.cpp:25-199): same module name, function names,
argument order, dtypes and shapes (int32 arrays, 65535 = unreachable, [x][y] order)."""
import numpy as np

from ctrm_b200 import _lib


def _as_int(a):
    """pybind11 `py::array_t<int>` force-casts; a float inf becomes INT_MIN on x86 (cost_to_go.cpp:112-113)"""
    a = np.asarray(a)
    if a.dtype.kind == "f":
        with np.errstate(invalid="ignore"):
            return np.ascontiguousarray(np.where(np.isfinite(a), a, -2147483648.0).astype(np.int64).astype(np.int32))
    return np.ascontiguousarray(a.astype(np.int32))


def getCostToGo2D(goal, occupancy_map, occupancy_agent):
    """creating cost_to_go"""
    lib = _lib.load()
    goal = np.ascontiguousarray(goal, dtype=np.float64)
    occ = _as_int(occupancy_map)
    agent = _as_int(occupancy_agent)
    M = occ.shape[0]
    out = np.empty((M, M), dtype=np.int32)
    _lib.check(lib.ctrm_host_cost_to_go_2d(_lib.ptr(goal), _lib.ptr(occ), M, _lib.ptr(agent), agent.shape[0], _lib.ptr(out)),
               "getCostToGo2D")
    return out


def getFov2dOccupancyCost(current_pos, occupancy_map, cost_to_go, fov_size, flatten):
    """get fov"""
    lib = _lib.load()
    pos = np.ascontiguousarray(current_pos, dtype=np.float64)
    occ = _as_int(occupancy_map)
    ctg = _as_int(cost_to_go)
    M, F = occ.shape[0], int(fov_size)
    out = np.empty(2 * F * F, dtype=np.int32)
    _lib.check(lib.ctrm_host_fov2d(_lib.ptr(pos), _lib.ptr(occ), _lib.ptr(ctg), M, F, _lib.ptr(out)), "getFov2dOccupancyCost")
    return out if flatten else out.reshape(2, F, F)
