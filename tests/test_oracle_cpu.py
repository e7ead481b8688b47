"""CPU tests (no GPU): the oracle against the golden vectors produced by the UNMODIFIED reference, the C restatement
against the numpy oracle, and the counter-based RNG known answers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ctrm_oracle as O
import rng_tables as R
from conftest import ROOT
from helpers import KeyedReplay, assert_same_roadmaps, golden_trm_arrays, oins_from_npz, otrms_to_arrays


@pytest.fixture(scope="module")
def oc():
    so = os.path.join(ROOT, "oracle", "liboracle_c.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle_c"], check=True, capture_output=True)
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ golden vectors of the reference
def test_sweep_golden(stage_gold, oc):
    z = stage_gold
    got = O.sweep_collide(z["s3.f1"], z["s3.t1"], z["s3.r1"], z["s3.c"], z["s3.c"], z["s3.r2"])
    assert np.array_equal(got, z["s3.verdict"])
    # C restatement, one obstacle per segment
    n = len(got)
    out = np.zeros(n, dtype=np.uint8)
    for s in range(0, n, 1):
        obs = np.array([[z["s3.c"][s, 0], z["s3.c"][s, 1], z["s3.r2"][s]]])
        oc.oc_collide_segments(C.c_int64(1), P(np.ascontiguousarray(z["s3.f1"][s])), P(np.ascontiguousarray(z["s3.t1"][s])),
                               P(np.ascontiguousarray(z["s3.r1"][s:s + 1])), 1, P(obs), P(out[s:s + 1]))
    assert np.array_equal(out.astype(bool), z["s3.verdict"])


def test_reference_collision_unit_cases():
    """tests/environment/test_fcl_utils.py:73-104 (2-D half; the oracle is 2-D like the path)"""
    a = lambda *v: np.array(v, dtype=np.float64)  # noqa: E731
    assert O.sweep_collide(a(0, 0), a(1, 1), 0.1, a(1, 0), a(0, 1), 0.1)
    assert not O.sweep_collide(a(0, 0), a(1, 0), 0.1, a(0, 1), a(1, 1), 0.1)


def test_occupancy_ctg_fov_golden(stage_gold, oc):
    z = stage_gold
    oi = oins_from_npz(z, "s1.")
    M, F = 160, 19
    occ = O.draw_occupancy(oi.obs_pos, oi.obs_rad, M)
    assert np.array_equal(occ, z["s1.occ"])
    occ_c = np.zeros((M, M), dtype=np.uint8)
    oc.oc_occupancy(M, len(oi.obs_rad), P(np.ascontiguousarray(np.concatenate([oi.obs_pos, oi.obs_rad[:, None]], 1))), P(occ_c))
    assert np.array_equal(occ_c, occ)
    for a in range(oi.num_agents):
        ctg = O.cost_to_go(oi.goals[a], occ, oi.rads[a])
        assert np.array_equal(ctg.astype(np.uint16), z["s1.ctg"][a])
        st = np.ascontiguousarray(O.agent_stencil(oi.rads[a], M).astype(np.int32))
        ctg_c = np.zeros((M, M), dtype=np.uint16)
        oc.oc_cost_to_go(M, P(occ), P(st), (st.shape[0] - 1) // 2, P(np.ascontiguousarray(oi.goals[a])), P(ctg_c))
        assert np.array_equal(ctg_c, z["s1.ctg"][a])
        fov = O.fov_crop_batch(z["s1.pts"], occ, np.stack([ctg] * len(z["s1.pts"])), F)
        assert np.array_equal(fov.astype(np.uint8), z["s1.fov"][a])
        out = np.zeros(2 * F * F, dtype=np.int32)
        for j, p in enumerate(z["s1.pts"][:8]):
            oc.oc_fov_crop(M, F, P(occ), P(ctg_c), P(np.ascontiguousarray(p)), P(out))
            assert np.array_equal(out.astype(np.uint8), z["s1.fov"][a, j])


def test_reference_fov_test_vectors(stage_gold):
    """tests/learning/test_fov.py:29-94 of the reference: cost 30 at the start cell, two 2x5x5 FOVs"""
    z = stage_gold
    obs_pos, obs_rad = np.array([[0.5, 0.5], [1.0, 0.0]]), np.array([0.2, 0.2])
    occ = O.draw_occupancy(obs_pos, obs_rad, 16)
    assert np.array_equal(occ, z["s4.occ"])
    ctg = O.cost_to_go(np.array([1.0, 1.0]), occ, 0.05)
    assert ctg[0, 0] == 30
    assert np.array_equal(ctg.astype(np.uint16), z["s4.ctg"])
    assert np.array_equal(O.fov_crop(np.array([0.0, 0.0]), occ, ctg, 5).astype(np.uint8), z["s4.fov00"].reshape(-1))
    assert np.array_equal(O.fov_crop(np.array([0.0, 1.0]), occ, ctg, 5).astype(np.uint8), z["s4.fov01"].reshape(-1))


def test_others_info_golden(stage_gold):
    z = stage_gold
    got = O.others_info(z["s2.cur"], z["s2.goals"], z["s2.prev"], z["s2.rads"], z["s2.speeds"])
    assert np.array_equal(got, z["s2.info"])


def test_networks_vs_reference_trace(trace_gold, oracle_weights):
    """get_feature_one_step and CTRMNet.sample at recorded steps of the reference run: bit-identical on this CPU build
    (same ATen ops in the same order) up to batch-shape dependent sgemm blocking: <= 2e-6 absolute"""
    import torch

    torch.set_num_threads(1)  # the reference run was recorded single-threaded
    z = trace_gold
    oi = oins_from_npz(z)
    occ, ctgs = O.set_instance(oracle_weights, oi)
    K = 3
    with torch.no_grad():
        for j, s in enumerate(z["nn_steps"]):
            X = O.feature_one_step(oracle_weights, oi, occ, ctgs, z["cur"][s], z["prev"][s])
            assert np.abs(X.numpy() - z["nn_X"][j]).max() <= 1e-6
            Xk = torch.cat([X] * K)
            IND, _ = O.indicator_onehot(oracle_weights, Xk)
            assert np.array_equal(IND.numpy().astype(np.int8), z["IND"][s])
            Y = O.cvae_sample(oracle_weights, Xk, IND, z["nn_U"][j]).numpy()
            assert np.abs(Y - z["SAMPLES"][s]).max() <= 2e-6


def test_reference_run_replay(any_trace_gold, oracle_weights):
    """the whole path: the oracle fed the reference's own random streams and sampled coordinates reproduces the
    reference's timed roadmaps and collision counter bit for bit"""
    z = any_trace_gold
    oi = oins_from_npz(z)
    NT, MT, MA = int(z["N_traj"]), int(z["max_T"]), int(z["max_attempt"])
    p = O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=float(z["merge_distance_rate"]), max_attempt=MA)
    steps = iter(range(len(z["SAMPLES"])))

    class StepTeacher(dict):
        def __getitem__(self, key):
            return z["SAMPLES"][next(steps)]

    rng = KeyedReplay(z["draws"], z["sincos"], None, oi.num_agents, NT, MT, MA)
    otrms = O.build_timed_roadmaps(oi, oracle_weights, p, rng, teacher=StepTeacher())
    assert_same_roadmaps(otrms_to_arrays(otrms), golden_trm_arrays(z))
    assert oi.cnt_continuous_collide == int(z["cnt_continuous_collide"])
    assert rng.i == len(z["draws"])  # every recorded draw consumed, in the reference's order


def test_timed_roadmap_container():
    """tests/roadmap/test_timed_roadmap.py:6-19 of the reference, on the oracle container"""
    trm = O.OTimedRoadmap(np.array([0.0, 0.0]))
    valid = lambda a, b: np.linalg.norm(a - b) <= 0.2  # noqa: E731
    for loc in (np.array([0.1, 0.0]), np.array([0.0, 0.1]), np.array([0.3, 0.0])):
        trm.append_sample_valid(loc, 1, valid)
    assert len(trm.V) == 2 and len(trm.V[1]) == 3
    assert trm.E[0][0] == [0, 1]
    trm.append_sample_valid(np.array([0.2, 0.0]), 2, valid)
    assert trm.E[1][0] == [0] and trm.E[1][1] == [] and trm.E[1][2] == [0]


# ------------------------------------------------------------------ counter-based RNG
def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10"""
    assert [int(x) for x in R.philox4x32(0, 0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(x) for x in R.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == [
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(x) for x in R.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)] == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_det_sincos_accuracy():
    u = np.random.RandomState(0).rand(2000)
    for v in list(u) + [0.0, 0.25, 0.5, 0.75, 0.125, 0.999999999]:
        s, c = R.det_sincos_2pi(float(v))
        assert abs(s - np.sin(2 * np.pi * v)) < 2e-15 and abs(c - np.cos(2 * np.pi * v)) < 2e-15  # np.sin(2*pi*v) itself rounds its argument


def test_philox_tables_ranges():
    t = R.PhiloxTables(123456789, 5)
    g = [t.gate(1, 2, a) for a in range(50)]
    assert all(0.0 <= x < 1.0 for x in g) and len(set(g)) == 50
    U = t.gumbel_uniforms(0, 1, 90, 64)
    assert U.shape == (90, 64) and U.dtype == np.float32 and U.min() >= 0.0 and U.max() < 1.0
    assert not np.array_equal(U, R.PhiloxTables(123456789, 6).gumbel_uniforms(0, 1, 90, 64))
