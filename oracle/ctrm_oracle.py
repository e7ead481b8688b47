"""CPU ORACLE — a restatement of the reference's timed-roadmap construction path.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product (ctrm_b200/) never does and fails loudly without its
CUDA library.  Every function cites the reference file:line it follows (paths relative to
/root/reference).  Integer/byte/fp64 geometry is numpy / plain Python floats (IEEE double, no FMA);
the fp32 networks use torch-CPU functional ops in the reference's own op order so that, fed the
reference's own random streams (rng_tables.StreamRNG), the oracle reproduces the reference bit for bit.

Pinning (oracle/gen_golden.py, tests/test_oracle_*.py):
  * the reference's own golden vectors: tests/learning/test_fov.py:29-94 (cost 30, two 2x5x5 FOVs),
    tests/environment/test_fcl_utils.py:73-104, tests/environment/test_static_objects.py:7-62
    (continuous half), tests/roadmap/test_timed_roadmap.py:6-19;
  * outputs of the unmodified reference run in the build container on notebooks/CTRM_demo_ins.pkl and on
    generated aamas22 instances, committed under tests/golden/ with the generating script.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

try:  # torch is only needed for the network stages
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None

UNREACHABLE = 65535


# --------------------------------------------------------------------------------------------
# a1  problem definition (environment/instance.py:17-46, environment/obstacle.py:40-48)
# --------------------------------------------------------------------------------------------
@dataclass
class OInstance:
    starts: np.ndarray      # (N,2) f64
    goals: np.ndarray       # (N,2) f64
    max_speeds: np.ndarray  # (N,) f64
    rads: np.ndarray        # (N,) f64
    obs_pos: np.ndarray     # (O,2) f64
    obs_rad: np.ndarray     # (O,) f64
    cnt_continuous_collide: int = 0  # static_objects.py:72-75

    @property
    def num_agents(self) -> int:
        return self.starts.shape[0]

    @staticmethod
    def from_arrays(starts, goals, max_speeds, rads, obs_pos, obs_rad) -> "OInstance":
        return OInstance(
            np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 2),
            np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, 2),
            np.ascontiguousarray(max_speeds, dtype=np.float64).reshape(-1),
            np.ascontiguousarray(rads, dtype=np.float64).reshape(-1),
            np.ascontiguousarray(obs_pos, dtype=np.float64).reshape(-1, 2),
            np.ascontiguousarray(obs_rad, dtype=np.float64).reshape(-1),
        )

    @staticmethod
    def from_reference(ins) -> "OInstance":
        """from a ctrm.environment.Instance-like object (reference or ctrm_b200 mirror)"""
        return OInstance.from_arrays(
            np.array(ins.starts), np.array(ins.goals), np.array(ins.max_speeds), np.array(ins.rads),
            np.array([o.pos for o in ins.obs]).reshape(-1, 2), np.array([o.rad for o in ins.obs]),
        )


# --------------------------------------------------------------------------------------------
# a2  occupancy raster (static_objects.py:173-187, obstacle.py:63-81, skimage.draw.disk 0.18.1)
# --------------------------------------------------------------------------------------------
def disk_cells(cx: float, cy: float, radius: float, shape):
    """skimage.draw.disk((cx,cy), radius, shape=shape) — see oracle/ref_stubs/skimage/draw.py"""
    center = np.array([cx, cy], dtype=np.float64)
    radii = np.array([radius, radius], dtype=np.float64)
    ul = np.ceil(center - radii).astype(int)
    lr = np.floor(center + radii).astype(int)
    ul = np.maximum(ul, 0)
    lr = np.minimum(lr, np.array(shape[:2]) - 1)
    bs = lr - ul + 1
    if bs[0] <= 0 or bs[1] <= 0:
        return np.zeros(0, dtype=np.intp), np.zeros(0, dtype=np.intp)
    r_lim, c_lim = np.ogrid[0:float(bs[0]), 0:float(bs[1])]
    sc = center - ul
    with np.errstate(divide="ignore", invalid="ignore"):
        d = ((r_lim - sc[0]) / radii[0]) ** 2 + ((c_lim - sc[1]) / radii[1]) ** 2
    rr, cc = np.nonzero(d < 1)
    return rr + ul[0], cc + ul[1]


def draw_occupancy(obs_pos: np.ndarray, obs_rad: np.ndarray, M: int) -> np.ndarray:
    """(M,M) uint8 {0,1}, first axis = x.  obstacle.py:73-79: X=int(M*x), Y=int(M*y), R=int(M*rad)."""
    occ = np.zeros((M, M), dtype=np.uint8)
    for p, r in zip(obs_pos, obs_rad):
        X = int(M * p[0])
        Y = int(M * p[1])
        R = int(M * r)
        rr, cc = disk_cells(X, Y, R, (M, M))
        occ[rr, cc] = 1
    return occ


# --------------------------------------------------------------------------------------------
# a3/a4  cost-to-go (learning/fov.py:34-63, cost_to_go.cpp:25-107)
# --------------------------------------------------------------------------------------------
def agent_stencil(rad: float, M: int) -> np.ndarray:
    """(2r+1,2r+1) int {0,1}: fov.py:48-56, r = ceil(rad*M)"""
    r = math.ceil(rad * M)
    st = np.zeros((2 * r + 1, 2 * r + 1), dtype=np.int32)
    rr, cc = disk_cells(r, r, r, st.shape)
    st[rr, cc] = 1
    return st


def grid_cell(v: float, M: int) -> int:
    """cost_to_go.cpp:41-46 / :121-126: v<1 ? floor(v*M) : M-1  (then wrapped through uint16)"""
    c = int(math.floor(v * M)) if v < 1 else M - 1
    return c & 0xFFFF


def passable_mask(occ: np.ndarray, stencil: np.ndarray) -> np.ndarray:
    """cost_to_go.cpp:68-97: cell is enterable iff occ==0 and no stencil cell with agent==1 lies outside
    the map or on an occupied cell."""
    M = occ.shape[0]
    r = (stencil.shape[0] - 1) // 2
    pad = np.ones((M + 2 * r, M + 2 * r), dtype=np.uint8)  # outside counts as blocked
    pad[r:r + M, r:r + M] = occ
    blocked = occ.astype(bool).copy()
    for k in range(-r, r + 1):
        for l in range(-r, r + 1):
            if stencil[k + r, l + r] == 1:
                blocked |= pad[r + k:r + k + M, r + l:r + l + M].astype(bool)
    return ~blocked


def cost_to_go(goal: np.ndarray, occ: np.ndarray, rad: float) -> np.ndarray:
    """getCostToGo2D (cost_to_go.cpp:25-107): unit-cost BFS from the goal cell through 4-connected
    passable cells; goal cell = 0 unconditionally (:50); unreachable = 65535; indexed [x][y]; int32."""
    M = occ.shape[0]
    ok = passable_mask(occ, agent_stencil(rad, M))
    gx, gy = grid_cell(goal[0], M), grid_cell(goal[1], M)
    ctg = np.full((M, M), UNREACHABLE, dtype=np.int32)
    if gx >= M or gy >= M:  # asserts compiled out in the reference; undefined there
        return ctg
    ctg[gx, gy] = 0
    frontier = np.zeros((M, M), dtype=bool)
    frontier[gx, gy] = True
    visited = frontier.copy()
    level = 0
    while frontier.any():
        level += 1
        nb = np.zeros_like(frontier)
        nb[1:, :] |= frontier[:-1, :]
        nb[:-1, :] |= frontier[1:, :]
        nb[:, 1:] |= frontier[:, :-1]
        nb[:, :-1] |= frontier[:, 1:]
        nb &= ok & ~visited
        ctg[nb] = level
        visited |= nb
        frontier = nb
    return ctg


# --------------------------------------------------------------------------------------------
# a5  FOV crop (cost_to_go.cpp:109-186)
# --------------------------------------------------------------------------------------------
def fov_crop(pos: np.ndarray, occ: np.ndarray, ctg: np.ndarray, F_: int) -> np.ndarray:
    """flatten=True layout: idx = x_fov*F + y_fov, channel 1 at +F*F; int32 {0,1}.
    ch0 = 1-occ; ch1 = reach(x,y) and (not reach(c) or ctg[x,y] < ctg[c]) where reach <=> != 65535
    (the reference sees inf cast to INT_MIN, i.e. `_c >= 0` is the reachability test)."""
    M = occ.shape[0]
    cx, cy = grid_cell(pos[0], M), grid_cell(pos[1], M)
    b = (F_ - 1) // 2
    out = np.zeros(2 * F_ * F_, dtype=np.int32)
    c = int(ctg[cx, cy])
    c_reach = c != UNREACHABLE
    for x in range(max(0, cx - b), min(M - 1, cx + b) + 1):
        for y in range(max(0, cy - b), min(M - 1, cy + b) + 1):
            i = (x - cx + b) * F_ + (y - cy + b)
            out[i] = 1 - int(occ[x, y])
            _c = int(ctg[x, y])
            out[i + F_ * F_] = int(_c != UNREACHABLE and ((not c_reach) or _c < c))
    return out


def fov_crop_batch(cur: np.ndarray, occ: np.ndarray, ctgs: np.ndarray, F_: int) -> np.ndarray:
    """vectorised fov_crop for all agents: (N, 2F^2) int32"""
    M = occ.shape[0]
    N = cur.shape[0]
    b = (F_ - 1) // 2
    out = np.zeros((N, 2, F_, F_), dtype=np.int32)
    for j in range(N):
        cx, cy = grid_cell(cur[j, 0], M), grid_cell(cur[j, 1], M)
        x0, x1 = max(0, cx - b), min(M - 1, cx + b)
        y0, y1 = max(0, cy - b), min(M - 1, cy + b)
        if x0 > x1 or y0 > y1:
            continue
        sub_o = occ[x0:x1 + 1, y0:y1 + 1]
        sub_c = ctgs[j][x0:x1 + 1, y0:y1 + 1]
        c = int(ctgs[j][cx, cy])
        reach = sub_c != UNREACHABLE
        ch1 = reach & ((c == UNREACHABLE) | (sub_c < c))
        out[j, 0, x0 - cx + b:x1 - cx + b + 1, y0 - cy + b:y1 - cy + b + 1] = 1 - sub_o
        out[j, 1, x0 - cx + b:x1 - cx + b + 1, y0 - cy + b:y1 - cy + b + 1] = ch1
    return out.reshape(N, 2 * F_ * F_)


# --------------------------------------------------------------------------------------------
# a8/a9  pairwise + self features (learning/compiled.py:11-125)
# --------------------------------------------------------------------------------------------
def normed_vec_mag(v: np.ndarray) -> np.ndarray:
    """compiled.py:11-27: (..,2) -> (..,3) [v/|v| (or v if |v|==0), |v|], float64"""
    mag = np.sqrt(np.sum(v * v, axis=-1, keepdims=True))
    safe = np.where(mag == 0, 1.0, mag)
    return np.concatenate([v / safe, mag], axis=-1)


def others_info(cur: np.ndarray, goals: np.ndarray, prev: np.ndarray, rads: np.ndarray,
                speeds: np.ndarray) -> np.ndarray:
    """compiled.py:30-97: row i*N+j = [nvm(cur_j-cur_i), nvm(goal_j-cur_i), nvm(prev_j-cur_i), rad_j, speed_j]
    (N*N, 11) float64"""
    N = cur.shape[0]
    ci = cur[:, None, :]
    out = np.empty((N, N, 11), dtype=np.float64)
    out[:, :, 0:3] = normed_vec_mag(cur[None, :, :] - ci)
    out[:, :, 3:6] = normed_vec_mag(goals[None, :, :] - ci)
    out[:, :, 6:9] = normed_vec_mag(prev[None, :, :] - ci)
    out[:, :, 9] = rads[None, :]
    out[:, :, 10] = speeds[None, :]
    return out.reshape(N * N, 11)


# --------------------------------------------------------------------------------------------
# weights (learning/utils.py:32-53, learning/model/__init__.py:19-59) as a flat dict of arrays
# --------------------------------------------------------------------------------------------
@dataclass
class OWeights:
    """state_dicts of the five networks + the hyper-parameters the path reads.
    keys: 'fov_self.mlp.0.weight', 'fov_vain...', 'agent.mlp.0.weight', 'model.indicator.0.weight', ..."""
    arrays: dict
    fov_size: int = 19
    map_size: int = 160
    num_neighbors: int = 15
    use_k_neighbor: bool = True
    include_self_attention: bool = False
    dim_ind: int = 3
    dim_latent: int = 64
    dim_message: int = 32
    dim_attention: int = 10
    temp: float = 2.0
    _t: dict = field(default_factory=dict)

    def t(self, key: str):
        if key not in self._t:
            self._t[key] = torch.from_numpy(np.ascontiguousarray(self.arrays[key]))
        return self._t[key]

    @staticmethod
    def load(path: str) -> "OWeights":
        z = np.load(path, allow_pickle=False)
        arrays = {k: z[k] for k in z.files if not k.startswith("hp.")}
        hp = {k[3:]: z[k].item() for k in z.files if k.startswith("hp.")}
        return OWeights(arrays=arrays, fov_size=int(hp["fov_size"]), map_size=int(hp["map_size"]),
                        num_neighbors=int(hp["num_neighbors"]), use_k_neighbor=bool(hp["use_k_neighbor"]),
                        include_self_attention=bool(hp["include_self_attention"]), dim_ind=int(hp["dim_ind"]),
                        dim_latent=int(hp["dim_latent"]), dim_message=int(hp["dim_message"]),
                        dim_attention=int(hp["dim_attention"]), temp=float(hp["temp"]))


# --------------------------------------------------------------------------------------------
# a7/a10/a11/a12  encoders + communication (fov_encoder.py:34-59, agent_encoder.py:34-65,
#                 format_input.py:212-277, :279-357)
# --------------------------------------------------------------------------------------------
def mlp3(w: OWeights, prefix: str, x):
    """Linear-ReLU-Linear-ReLU-Linear, state-dict keys mlp.0/.2/.4 (no batch norm in the trained config)"""
    h = F.relu(F.linear(x, w.t(prefix + ".mlp.0.weight"), w.t(prefix + ".mlp.0.bias")))
    h = F.relu(F.linear(h, w.t(prefix + ".mlp.2.weight"), w.t(prefix + ".mlp.2.bias")))
    return F.linear(h, w.t(prefix + ".mlp.4.weight"), w.t(prefix + ".mlp.4.bias"))


def mlpbn(w: OWeights, prefix: str, x):
    """cvae.py:39-55 in eval mode: Linear, BN, ReLU, Linear, BN, ReLU, Linear (keys .0 .1 .3 .4 .6)"""
    def bn(h, p):
        return F.batch_norm(h, w.t(p + ".running_mean"), w.t(p + ".running_var"), w.t(p + ".weight"),
                            w.t(p + ".bias"), False, 0.1, 1e-5)
    h = F.relu(bn(F.linear(x, w.t(prefix + ".0.weight"), w.t(prefix + ".0.bias")), prefix + ".1"))
    h = F.relu(bn(F.linear(h, w.t(prefix + ".3.weight"), w.t(prefix + ".3.bias")), prefix + ".4"))
    return F.linear(h, w.t(prefix + ".6.weight"), w.t(prefix + ".6.bias"))


def get_comm(num_agents: int, num_neighbors: int, info, attentions, messages, include_self_attention: bool):
    """format_input.py:212-277 with T=1 (same torch ops in the same order)"""
    T = 1
    neighbors_idx = torch.argsort(info.reshape(T, num_agents, num_agents, -1)[:, :, :, 2]).reshape(
        T, num_agents, num_agents, 1)
    att_sorted = torch.gather(attentions.reshape(T, num_agents, num_agents, -1), 2,
                              neighbors_idx.expand(T, num_agents, num_agents, attentions.shape[-1])
                              )[:, :, : num_neighbors + 1]
    msg_sorted = torch.gather(messages.reshape(T, num_agents, num_agents, -1), 2,
                              neighbors_idx.expand(T, num_agents, num_agents, messages.shape[-1])
                              )[:, :, 1: num_neighbors + 1]
    diff = att_sorted - att_sorted[:, :, 0].reshape(T, num_agents, 1, -1)
    sim = -torch.sum(torch.pow(diff, 2), dim=-1)
    if not include_self_attention:
        weight = F.softmax(sim[:, :, 1:], dim=-1)
    else:
        weight = F.softmax(sim, dim=-1)[:, :, 1:]
    return torch.sum(msg_sorted * weight.reshape(T, num_agents, num_neighbors, 1), dim=2).reshape(num_agents, -1)


def feature_one_step(w: OWeights, ins: OInstance, occ, ctgs, cur: np.ndarray, prev: np.ndarray, rec=None):
    """Format2D_CTRM_Input.get_feature_one_step (format_input.py:279-357) -> (N,72) float32 tensor"""
    N = ins.num_agents
    fov = fov_crop_batch(cur, occ, ctgs, w.fov_size)
    x_fov = torch.from_numpy(fov).float()
    e_vain = mlp3(w, "fov_vain", x_fov)
    e_self = mlp3(w, "fov_self", x_fov)
    info64 = others_info(cur, ins.goals, prev, ins.rads, ins.max_speeds)
    info = torch.from_numpy(info64).float()
    x_vain = torch.cat([info, torch.cat([e_vain] * N)], dim=1)
    y = mlp3(w, "agent", x_vain)
    messages, attentions = torch.split(y, [w.dim_message, w.dim_attention], dim=-1)
    k = N - 1
    if w.use_k_neighbor:
        k = min(w.num_neighbors, k)
    comm = get_comm(N, k, info, attentions, messages, w.include_self_attention)
    idx = torch.arange(N) * N + torch.arange(N)
    self_info = info[:, 3:][idx]
    X = torch.cat([self_info, e_self, comm], dim=1)
    if rec is not None:
        rec["fov"] = fov
        rec["e_self"] = e_self.numpy().copy()
        rec["e_vain"] = e_vain.numpy().copy()
        rec["info"] = info.numpy().copy()
        rec["messages"] = messages.numpy().copy()
        rec["attentions"] = attentions.numpy().copy()
        rec["comm"] = comm.numpy().copy()
        rec["X"] = X.numpy().copy()
    return X


# --------------------------------------------------------------------------------------------
# a13/a14  indicator + CVAE sampling (cvae.py:77-79, :128-138; torch RelaxedOneHotCategorical)
# --------------------------------------------------------------------------------------------
def indicator_onehot(w: OWeights, Xk):
    """learned_sampler.py:99-104"""
    logits = mlpbn(w, "model.indicator", Xk)
    return F.one_hot(torch.argmax(logits, -1), w.dim_ind), logits


def cvae_sample(w: OWeights, Xk, IND, uniforms: np.ndarray, rec=None):
    """CTRMNet.sample (cvae.py:128-138) with the uniforms of RelaxedOneHotCategorical.rsample injected.
    torch: probs = exp(logp); Categorical normalises probs/sum, logits = log(clamp(probs, eps, 1-eps));
    u = clamp(U, eps, 1-eps); g = -log(-log u); s = (logits+g)/temp; z = exp(s - logsumexp(s))."""
    x = torch.cat((Xk, IND), -1)
    log_prob_x = F.log_softmax(mlpbn(w, "model.encoder_input", x), dim=-1)
    probs = torch.exp(log_prob_x)
    probs = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(torch.float32).eps
    logits = torch.log(probs.clamp(min=eps, max=1 - eps))
    u = torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float32)).clamp(min=eps, max=1 - eps)
    gumbels = -((-(u.log())).log())
    scores = (logits + gumbels) / w.temp
    z = (scores - scores.logsumexp(dim=-1, keepdim=True)).exp()
    y = mlpbn(w, "model.decoder", torch.cat([z, x], -1))
    if rec is not None:
        rec["log_prob_x"] = log_prob_x.numpy().copy()
        rec["z"] = z.numpy().copy()
    return y


# --------------------------------------------------------------------------------------------
# a16/a17  validity + swept collision (roadmap/utils.py:18-42, static_objects.py:108-138,
#          collision_check.cpp:16-58)
# --------------------------------------------------------------------------------------------
def sweep_collide(f1: np.ndarray, t1: np.ndarray, rad1: float, f2: np.ndarray, t2: np.ndarray, rad2) -> np.ndarray:
    """continuousCollideSpheres vectorised over leading axes; last axis = dim.  Operation order is the
    reference's (collision_check.cpp:33-35, :41-42); numpy ufuncs round every * and + (no FMA)."""
    val = ((-f1) + t1 + f2) - t2
    a = np.zeros(val.shape[:-1])
    b = np.zeros(val.shape[:-1])
    d12 = f1 - f2
    for i in range(val.shape[-1]):
        a = a + val[..., i] * val[..., i]
        b = b + val[..., i] * d12[..., i]
    with np.errstate(divide="ignore", invalid="ignore"):
        tt = np.where(b >= 0, 0.0, np.where(a + b <= 0, 1.0, -b / a))
    tt = tt[..., None]
    v = ((1 - tt) * f1 + tt * t1) - ((1 - tt) * f2 + tt * t2)
    tot = np.zeros(val.shape[:-1])
    for i in range(val.shape[-1]):
        tot = tot + v[..., i] * v[..., i]
    rad = rad1 + rad2
    return tot <= rad * rad


def collide_continuous_sphere(ins: OInstance, p1: np.ndarray, p2: np.ndarray, rad: float) -> bool:
    """StaticObjects.collide_continuous_sphere (static_objects.py:59-75, :108-138): any over obstacles"""
    ins.cnt_continuous_collide += 1
    if ins.obs_pos.shape[0] == 0:
        return False
    return bool(np.any(sweep_collide(p1[None, :], p2[None, :], rad, ins.obs_pos, ins.obs_pos, ins.obs_rad)))


def valid_move(ins: OInstance, p1: np.ndarray, p2: np.ndarray, speed: float, rad: float) -> bool:
    """roadmap/utils.py:18-42"""
    speed = float(speed)  # Instance.max_speeds holds Python floats: `max_speed ** 2` is libm pow
    p = np.abs(p1 - p2)
    if p[0] > speed or p[1] > speed:
        return False
    p2_ = p * p  # ndarray ** 2 is np.square (x*x)
    if (0 + p2_[0]) + p2_[1] > speed ** 2:
        return False
    return not collide_continuous_sphere(ins, p1, p2, rad)


def valid_edge(ins: OInstance, i: int, p1: np.ndarray, p2: np.ndarray) -> bool:
    """closure of learned_sampler.py:127-132"""
    rad = float(ins.rads[i])
    if min(p2) < rad or 1 - rad < max(p2):
        return False
    return valid_move(ins, p1, p2, float(ins.max_speeds[i]), rad)


# --------------------------------------------------------------------------------------------
# a24  TimedRoadmap container (roadmap/timed_roadmap.py:45-150, :212-272)
# --------------------------------------------------------------------------------------------
class OTimedRoadmap:
    def __init__(self, start: np.ndarray):
        self.V: list = [[start]]       # V[t][i] = pos (f64[2]);  index == position in list
        self.E: list = [[[]]]          # E[t][i] = list of indices into V[t+1]

    def extend_layer(self, t: int):
        while t > len(self.V) - 1:
            self.V.append([])
            self.E.append([])

    def get_parents(self, t: int, idx: int):
        """timed_roadmap.py:212-233"""
        if t < 1:
            return []
        return [j for j in range(len(self.V[t - 1])) if idx in self.E[t - 1][j]]

    def append_sample(self, loc, t, parents, children):
        """timed_roadmap.py:67-104 with explicit parents/children"""
        self.extend_layer(t)
        idx = len(self.V[t])
        self.V[t].append(loc)
        self.E[t].append([])
        for p in parents:
            self.E[t - 1][p].append(idx)
        for c in children:
            self.E[t][idx].append(c)

    def append_sample_valid(self, loc, t, valid):
        """timed_roadmap.py:67-150 with valid_edge: add_parents then add_children"""
        self.extend_layer(t)
        idx = len(self.V[t])
        self.V[t].append(loc)
        self.E[t].append([])
        for i, u in enumerate(self.V[t - 1]):
            if valid(u, loc):
                self.E[t - 1][i].append(idx)
        if not (t + 1 > len(self.V) - 1):
            for i, u in enumerate(self.V[t + 1]):
                if valid(loc, u):
                    self.E[t][idx].append(i)

    def extend_until(self, T: int, valid):
        """timed_roadmap.py:248-272"""
        n = len(self.V[-1])
        adjs = [[i] for i in range(n)]
        for i in range(n):
            for j in range(i + 1, n):
                if valid(self.V[-1][i], self.V[-1][j]):
                    adjs[i].append(j)
                    adjs[j].append(i)
        while len(self.V) - 1 < T:
            self.V.append([p for p in self.V[-1]])
            self.E[-1] = [list(a) for a in adjs]
            self.E.append([[] for _ in range(n)])

    def n_vertices(self) -> int:
        return sum(len(v) for v in self.V)

    def n_edges(self) -> int:
        return sum(len(e) for layer in self.E for e in layer)


# --------------------------------------------------------------------------------------------
# a20  merge_samples (roadmap_learned/utils.py:16-137)
# --------------------------------------------------------------------------------------------
def merge_samples(loc: np.ndarray, t: int, agent: int, trm: OTimedRoadmap, ins: OInstance,
                  merge_distance: float) -> np.ndarray:
    rad = float(ins.rads[agent])
    speed = float(ins.max_speeds[agent])
    goal = ins.goals[agent]
    has_next = t + 1 <= len(trm.V) - 1
    has_t = len(trm.V) > t

    def d2(cands):
        c = np.array(cands).reshape(-1, 2)
        d = c - loc
        d = d * d
        return (0 + d[:, 0]) + d[:, 1]

    dp = d2(trm.V[t - 1])
    parents = [int(i) for i in np.where(dp <= speed ** 2)[0]
               if not collide_continuous_sphere(ins, trm.V[t - 1][i], loc, rad)]
    set_p = set(parents)
    if has_next:
        dc = d2(trm.V[t + 1])
        children = [int(i) for i in np.where(dc <= speed ** 2)[0]
                    if not collide_continuous_sphere(ins, trm.V[t + 1][i], loc, rad)]
    else:
        children = []
    set_c = set(children)

    if has_t:
        dm = d2(trm.V[t])
        h_loc = sum((loc - goal) * (loc - goal))
        for u in np.where(dm <= merge_distance ** 2)[0]:
            u = int(u)
            up = set(trm.get_parents(t, u))
            uc = set(trm.E[t][u])
            if up == set_p and uc == set_c:
                h_u = sum((trm.V[t][u] - goal) * (trm.V[t][u] - goal))
                if h_loc < h_u:
                    trm.V[t][u] = loc
                    return loc
                return trm.V[t][u]
            if up >= set_p and uc >= set_c:
                return trm.V[t][u]
            if up <= set_p and uc <= set_c:
                trm.V[t][u] = loc
                trm.E[t][u] += list(set_c - uc)
                for p in set_p - up:
                    trm.E[t - 1][p].append(u)
                return loc
    trm.append_sample(loc, t, parents, children)
    return loc


# --------------------------------------------------------------------------------------------
# a22/a23  finalisation (roadmap_learned/utils.py:140-174)
# --------------------------------------------------------------------------------------------
def format_trms(ins: OInstance, trms):
    T = max(len(trm.V) for trm in trms) - 1
    for i, trm in enumerate(trms):
        trm.extend_until(T + 1, lambda a, b, i=i: valid_move(ins, a, b, float(ins.max_speeds[i]), float(ins.rads[i])))


def append_goals(ins: OInstance, trms):
    for i, trm in enumerate(trms):
        goal = ins.goals[i]
        for t in range(1, len(trm.V)):
            trm.append_sample_valid(goal, t, lambda a, b, i=i: valid_move(ins, a, b, float(ins.max_speeds[i]), float(ins.rads[i])))


# --------------------------------------------------------------------------------------------
# the roll-out driver (roadmap_learned/learned_sampler.py:19-204)
# --------------------------------------------------------------------------------------------
@dataclass
class OParams:
    prob_uniform_sampling_after_goal: float = 0.9
    prob_uniform_bias: float = 0.0
    prob_uniform_gamma: float = 5.0
    max_T: int = 64
    N_traj: int = 25
    max_attempt: int = 3
    randomize_indicator: bool = False
    merge_distance_rate: float = 0.25
    K: Optional[int] = None  # candidates per agent-step; None -> dim_ind (the reference couples them)


def prob_random_walk(p: OParams, t: int, makespan) -> float:
    """learned_sampler.py:84-91"""
    return float((1 - np.exp(-p.prob_uniform_gamma * t / (p.max_T if makespan is None else makespan)))
                 * (1 - p.prob_uniform_bias) + p.prob_uniform_bias)


def set_instance(w: OWeights, ins: OInstance):
    """Format2D_CTRM_Input.set_instance (format_input.py:200-210)"""
    occ = draw_occupancy(ins.obs_pos, ins.obs_rad, w.map_size)
    ctgs = np.stack([cost_to_go(ins.goals[i], occ, ins.rads[i]) for i in range(ins.num_agents)])
    return occ, ctgs


def build_timed_roadmaps(ins: OInstance, w: OWeights, p: OParams, rng, recorder=None, teacher=None):
    """get_timed_roadmaps_multiple_paths_with_learned_indicator (learned_sampler.py:19-204).

    recorder: optional list; one dict per executed step with every intermediate (for stage-wise parity).
    teacher : optional dict (k_traj,t) -> SAMPLES (K*N,3) f32 to use INSTEAD of the network output
              (teacher forcing: "given the reference's sampled coordinates", BASELINE.json north_star).
    Returns list[OTimedRoadmap]."""
    torch.set_num_threads(1)
    N = ins.num_agents
    K = p.K if p.K is not None else w.dim_ind
    occ, ctgs = set_instance(w, ins)
    trms = [OTimedRoadmap(ins.starts[i].copy()) for i in range(N)]
    makespan = None
    with torch.no_grad():
        for k_traj in range(p.N_traj):
            nxt: list = []
            cur = ins.starts.copy()
            prev = ins.starts.copy()
            arrived = [False] * N
            for t in range(1, p.max_T + 1):
                rec = {} if recorder is not None else None
                X = feature_one_step(w, ins, occ, ctgs, cur, prev, rec)
                Xk = torch.cat([X] * K)
                p_rw = prob_random_walk(p, t, makespan)
                u_step = rng.step_draw(k_traj, t)  # ALWAYS drawn (non-short-circuit `&`, :93)
                if p.randomize_indicator and (u_step > p_rw):
                    IND = F.one_hot(torch.from_numpy(rng.rand_indicator(k_traj, t, N * K, w.dim_ind)), w.dim_ind)
                    ind_logits = None
                else:
                    IND, ind_logits = indicator_onehot(w, Xk)
                U = rng.gumbel_uniforms(k_traj, t, K * N, w.dim_latent)
                SAMPLES = cvae_sample(w, Xk, IND, U, rec).numpy()
                if rec is not None:
                    rec.update(k_traj=k_traj, t=t, cur=cur.copy(), prev=prev.copy(), arrived=list(arrived),
                               makespan=makespan, p_rw=p_rw, U=np.array(U).copy(), IND=IND.numpy().copy(),
                               ind_logits=None if ind_logits is None else ind_logits.numpy().copy(),
                               SAMPLES_nn=SAMPLES.copy())
                if teacher is not None:
                    SAMPLES = np.asarray(teacher[(k_traj, t)], dtype=np.float32)
                # learned_sampler.py:114-121 (f32 product, promoted f64 add)
                cur_rep = np.repeat(cur[None], K, axis=0).reshape(-1, 2)
                Y = cur_rep + SAMPLES[:, :2] * SAMPLES[:, 2].reshape(-1, 1)
                if rec is not None:
                    rec.update(SAMPLES=SAMPLES.copy(), Y=Y.copy(), loc_pred=[], loc_next=[], branch=[])
                for i in range(N):
                    loc_cur = cur[i]
                    speed = float(ins.max_speeds[i])

                    def sample_uniform_i():
                        for att in range(p.max_attempt):
                            u_mag, s, c = rng.random_walk(k_traj, t, i, att)
                            mag = speed * u_mag
                            loc = np.array([s, c]) * mag + loc_cur
                            if valid_edge(ins, i, loc_cur, loc):
                                return loc
                        return loc_cur

                    u_gate = rng.gate(k_traj, t, i)
                    if ((not arrived[i] and u_gate < p_rw)
                            or (arrived[i] and u_gate > p.prob_uniform_sampling_after_goal)
                            or (k_traj < w.dim_ind)):
                        ok = False
                        for k in range(K):
                            loc_pred = Y[i + k * N]
                            if valid_edge(ins, i, loc_cur, loc_pred):
                                ok = True
                                break
                        branch = 1 + k if ok else 0
                        if not ok:
                            loc_pred = sample_uniform_i()
                    else:
                        branch = -1
                        loc_pred = sample_uniform_i()
                    loc_next = merge_samples(loc_pred, t, i, trms[i], ins, p.merge_distance_rate * speed)
                    nxt.append(loc_next)
                    if valid_edge(ins, i, loc_next, ins.goals[i]):
                        arrived[i] = True
                    if rec is not None:
                        rec["loc_pred"].append(np.array(loc_pred).copy())
                        rec["loc_next"].append(np.array(loc_next).copy())
                        rec["branch"].append(branch)
                if rec is not None:
                    rec["arrived_after"] = list(arrived)
                    recorder.append(rec)
                if all(arrived):
                    makespan = t if makespan is None else max(makespan, t)
                    break
                prev = cur
                cur = np.array(nxt)
                nxt = []
    format_trms(ins, trms)
    append_goals(ins, trms)
    return trms


def trm_counts(trms):
    return sum(t.n_vertices() for t in trms), sum(t.n_edges() for t in trms)
