"""Training-time input features of a demonstration on the GPU — the x half of Dataset.map_instance_to_tensor
(learning/dataset.py:113-160): for every time step t < T and agent, the FOV tensor
(getFov2dOccupancyCost, cost_to_go.cpp:109-186) around the agent's position on its demonstrated path, the pairwise
others_info rows (numba get_arr_others_info, learning/compiled.py:30-97) and self_info (compiled.py:100-125).
A straight batched reuse of the stage kernels: occupancy raster, cost-to-go wavefront, FOV raster, plus one pairwise kernel.
No CPU fallback."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import _lib


def demonstration_features(ins, path_pos: np.ndarray, T: Optional[int] = None, map_size: int = 160, fov_size: int = 19,
                           device: int = 0, to_host: bool = True) -> dict:
    """ins: Instance; path_pos [N][>=T+1][2] float64, the solution paths (e.g. PrioritizedPlanning.path_pos);
    T = number of time steps to featurise (default: all moves, len - 1, the reference's int(res.maximum_costs) is the caller's).
    Returns fovs (T, N, 2 F^2) f32, others_info (T, N*N, 11) f32, self_info (T, N, 8) f32 — numpy arrays, or CUDA tensors
    with to_host=False."""
    import torch

    lib = _lib.load()
    dev = torch.device("cuda", device)
    starts, goals, speeds, rads, obs = ins.arrays()
    N = len(starts)
    path_pos = np.ascontiguousarray(path_pos, dtype=np.float64)
    if path_pos.ndim != 3 or path_pos.shape[0] != N or path_pos.shape[2] != 2:
        raise _lib.CtrmError("demonstration_features: path_pos must be [num_agents][steps + 1][2]")
    if T is None:
        T = path_pos.shape[1] - 1
    if T < 1 or T + 1 > path_pos.shape[1]:
        raise _lib.CtrmError("demonstration_features: T out of range")
    M, F = int(map_size), int(fov_size)
    st = torch.cuda.current_stream(dev).cuda_stream
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    # cur[t][i] = path[i][t], prev[t][i] = path[i][max(0, t-1)]   (dataset.py:141-147)
    cur = np.ascontiguousarray(path_pos[:, :T].transpose(1, 0, 2))
    prev = np.ascontiguousarray(path_pos[:, np.maximum(np.arange(T) - 1, 0)].transpose(1, 0, 2))
    d_cur, d_prev = up(cur, np.float64), up(prev, np.float64)
    d_goals, d_rads, d_speeds = up(goals, np.float64), up(rads, np.float64), up(speeds, np.float64)
    # occupancy + one cost-to-go field per agent (dataset.py:119-127)
    d_obs = up(obs.reshape(-1, 3) if len(obs) else np.zeros((1, 3)), np.float64)
    d_off = up(np.array([0, len(obs)]), np.int32)
    occ = torch.zeros(M * M + 16, dtype=torch.uint8, device=dev)
    _lib.check(lib.ctrm_raster_occupancy(d_obs.data_ptr(), d_off.data_ptr(), 1, M, occ.data_ptr(), st), "ctrm_raster_occupancy")
    inst = torch.zeros(N, dtype=torch.int32, device=dev)
    sr = up(np.array([math.ceil(float(r) * M) for r in rads]), np.int32)
    ctg = torch.empty(N * M * M, dtype=torch.int16, device=dev)
    _lib.check(lib.ctrm_cost_to_go(occ.data_ptr(), inst.data_ptr(), d_goals.data_ptr(), sr.data_ptr(), N, M, ctg.data_ptr(), st),
               "ctrm_cost_to_go")
    # FOV tensor per time step (dataset.py:128-137): one launch per t, every agent reads its own field
    n_pad = N + (N & 1)  # 2 F^2 floats per row: an even row count keeps every time slice 16-byte aligned for the kernel's float4 stores
    fovs = torch.empty((T, n_pad, 2 * F * F), dtype=torch.float32, device=dev)
    for t in range(T):
        _lib.check(lib.ctrm_fov_raster(occ.data_ptr(), ctg.data_ptr(), 0, inst.data_ptr(), d_cur[t].data_ptr(), N, M, F,
                                       fovs[t].data_ptr(), st), "ctrm_fov_raster")
    info = torch.empty((T, N * N, 11), dtype=torch.float32, device=dev)
    self_info = torch.empty((T, N, 8), dtype=torch.float32, device=dev)
    _lib.check(lib.ctrm_others_info(d_cur.data_ptr(), d_goals.data_ptr(), d_prev.data_ptr(), d_rads.data_ptr(), d_speeds.data_ptr(),
                                    T, N, info.data_ptr(), self_info.data_ptr(), st), "ctrm_others_info")
    torch.cuda.current_stream(dev).synchronize()
    out = dict(fovs=fovs[:, :N].contiguous(), others_info=info, self_info=self_info)
    return {k: v.cpu().numpy() for k, v in out.items()} if to_host else out
