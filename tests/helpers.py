"""shared test helpers: oracle instances, keyed RNG tables from a recorded stream, TRM comparison"""
import numpy as np

import ctrm_oracle as O
from rng_tables import ReplayRNG

from ctrm_b200.environment import Instance


def oins_from_npz(z, prefix=""):
    return O.OInstance.from_arrays(z[prefix + "starts"], z[prefix + "goals"], z[prefix + "max_speeds"], z[prefix + "rads"],
                                   z[prefix + "obs_pos"], z[prefix + "obs_rad"])


def instance_from_oins(oi) -> Instance:
    return Instance.from_arrays(oi.starts, oi.goals, oi.max_speeds, oi.rads, oi.obs_pos, oi.obs_rad)


class KeyedReplay(ReplayRNG):
    """ReplayRNG that also records which key (k_traj, t, agent, attempt) consumed which draw, so a stream-order
    recording of the reference can be fed to the GPU as explicit keyed tables (ctrm_tables)."""

    def __init__(self, draws, sincos, uniforms, N, N_traj, max_T, max_attempt):
        super().__init__(draws, sincos, uniforms)
        self.t_gate = np.full((N_traj, max_T, N), np.nan)
        self.t_rw = np.full((N_traj, max_T, N, max_attempt, 3), np.nan)

    def gate(self, k_traj, t, agent):
        v = super().gate(k_traj, t, agent)
        self.t_gate[k_traj, t - 1, agent] = v
        return v

    def random_walk(self, k_traj, t, agent, attempt):
        u, s, c = super().random_walk(k_traj, t, agent, attempt)
        self.t_rw[k_traj, t - 1, agent, attempt] = (u, s, c)
        return u, s, c


def teacher_table(recs, N, K, N_traj, max_T, key="SAMPLES"):
    """oracle recorder -> [N_traj][max_T][N][K][3] float32 (reference row k*N+i -> [i][k])"""
    tab = np.zeros((N_traj, max_T, N, K, 3), dtype=np.float32)
    for r in recs:
        s = np.asarray(r[key], dtype=np.float32).reshape(K, N, 3)
        tab[r["k_traj"], r["t"] - 1] = s.transpose(1, 0, 2)
    return tab


def otrms_to_arrays(trms):
    """list[OTimedRoadmap] -> (nlayers, layer_ptr, pos, eptr, eidx sorted per vertex)"""
    layer_ptr, pos, eptr, eidx, nlayers = [0], [], [0], [], []
    for trm in trms:
        nlayers.append(len(trm.V))
        for t in range(len(trm.V)):
            for i, v in enumerate(trm.V[t]):
                pos.append(np.asarray(v, dtype=np.float64))
                eidx.extend(sorted(int(e) for e in trm.E[t][i]))
                eptr.append(len(eidx))
            layer_ptr.append(len(pos))
    return (np.array(nlayers), np.array(layer_ptr), np.array(pos).reshape(-1, 2), np.array(eptr), np.array(eidx))


def golden_trm_arrays(z, prefix="trm."):
    """fixture arrays with each adjacency list sorted (edge SETS are the parity target)"""
    eptr, eidx = z[prefix + "eptr"], z[prefix + "eidx"].copy()
    for v in range(len(eptr) - 1):
        eidx[eptr[v]:eptr[v + 1]] = np.sort(eidx[eptr[v]:eptr[v + 1]])
    return z[prefix + "nlayers"], z[prefix + "layer_ptr"], z[prefix + "pos"], eptr, eidx


def csr_to_arrays(csr, b=0):
    """CSRRoadmaps instance b -> same tuple as otrms_to_arrays (layer_ptr only over existing layers)"""
    a0, a1 = int(csr.agent_off[b]), int(csr.agent_off[b + 1])
    nlayers = csr.n_layers[a0:a1].astype(np.int64)
    lp = [0]
    v_first = int(csr.layer_ptr[a0 * csr.L])
    for a in range(a0, a1):
        for t in range(int(csr.n_layers[a])):
            lp.append(int(csr.layer_ptr[a * csr.L + t + 1]) - v_first)
    v_last = int(csr.layer_ptr[a1 * csr.L])
    pos = csr.pos[v_first:v_last]
    eptr = csr.eptr[v_first:v_last + 1] - csr.eptr[v_first]
    eidx = csr.eidx[int(csr.eptr[v_first]):int(csr.eptr[v_last])]
    return nlayers, np.array(lp), pos, eptr, eidx


def assert_same_roadmaps(got, want):
    names = ("nlayers", "layer_ptr", "pos", "eptr", "eidx")
    for n, g, w in zip(names, got, want):
        g, w = np.asarray(g), np.asarray(w)
        assert g.shape == w.shape, f"{n}: shape {g.shape} vs {w.shape}"
        if n == "pos":
            assert np.array_equal(g.view(np.uint64), w.view(np.uint64)), f"{n}: positions differ bitwise"
        else:
            assert np.array_equal(g, w), f"{n} differs"


def rel_err(a, ref):
    """|a - ref| / max(|ref|, rms(ref)) — the 1e-5 'relative' criterion of BASELINE.json with a scale floor
    (without a floor even two correct CPU fp32 implementations exceed 1e-5 on near-zero entries)"""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    floor = np.sqrt(np.mean(ref * ref)) if ref.size else 1.0
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0
