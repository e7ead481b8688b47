"""Teacher-forced end-to-end parity on instance shapes beyond the demo: more candidates than indicators (K != dim_ind),
more roll-outs than one 32-bit adjacency word (N_traj > 31), hetero radii / speeds, dense and absent obstacles, a single
agent, a larger crowd and a different grid resolution.  Topology, positions and the collision counter must equal the
oracle's bit for bit; the GPU's own CVAE outputs must stay within 1e-5 relative at every executed step."""
import numpy as np
import pytest

import ctrm_oracle as O
from conftest import WEIGHTS
from helpers import assert_same_roadmaps, csr_to_arrays, otrms_to_arrays, rel_err, teacher_table
from rng_tables import PhiloxTables

pytestmark = pytest.mark.gpu


def _instance(seed, n_lo, n_hi, obs_num, speeds="0.03125", rads="0.015625", obs_lo=0.05, obs_hi=0.08):
    from ctrm_b200.environment import generate_ins_2d_with_obs_hetero_nonfix_agents as gen

    return gen(n_lo, n_hi, speeds, rads, obs_num, obs_lo, obs_hi, rng=np.random.RandomState(seed))


def _oins(ins):
    s, g, sp, r, o = ins.arrays()
    return O.OInstance.from_arrays(s, g, sp, r, o[:, :2], o[:, 2])


CASES = [
    # name,            instances (seed, n_lo, n_hi, obs, speeds, rads),                      N_traj, max_T, K, rate, M
    ("k5",             [(1, 12, 13, 6, "0.03125", "0.015625")],                               2, 40, 5, 0.1, 160),
    ("ntraj40_w2",     [(2, 5, 6, 4, "0.03125", "0.015625")],                                 40, 12, 3, 0.25, 160),
    ("hetero",         [(3, 10, 11, 8, "0.03125,0.046875", "0.015625,0.0234375,0.01")],       2, 48, 3, 0.1, 160),
    ("dense_and_none", [(4, 8, 9, 20, "0.03125", "0.015625"), (5, 7, 8, 0, "0.03125", "0.015625")], 2, 48, 3, 0.1, 160),
    ("single_agent",   [(6, 1, 2, 3, "0.03125", "0.015625"), (7, 3, 4, 3, "0.03125", "0.015625")], 3, 32, 3, 0.1, 160),
    ("crowd60",        [(8, 60, 61, 10, "0.03125", "0.0078125")],                             1, 24, 3, 0.1, 160),
    ("map250_k4",      [(9, 9, 10, 6, "0.03125", "0.015625")],                                2, 40, 4, 0.1, 250),
    # more obstacles than the step kernel stages in shared memory (global-memory fallback), and the widest adjacency mask
    ("obs70",          [(10, 6, 7, 70, "0.03125", "0.0078125")],                              2, 32, 3, 0.1, 160),
    ("ntraj100_w4",    [(11, 3, 4, 3, "0.03125", "0.015625")],                                100, 8, 3, 0.25, 160),
    # BASELINE configs[1] at its real roll-out sizes: one benchmark instance, every one of its ~1,400 steps compared
    ("full_size",      [(46, 21, 30, 10, "0.03125", "0.015625")],                             25, 64, 3, 0.1, 160),
    # BASELINE configs[2] ("homo-more-agents", scripts/exp_scripts/benchmark_gen_hetero.sh:31): 31-40 agents at the full
    # N_traj = 25 x max_T = 64, every executed step compared
    ("more_agents_full", [(47, 31, 40, 10, "0.03125", "0.015625")],                           25, 64, 3, 0.1, 160),
    # BASELINE configs[4], the stress shape (100 agents, 64 obstacles, K = 64 candidates, map_size 250) on a short horizon
    ("stress_short",   [(900, 100, 101, 64, "0.03125", "0.00390625", 0.02, 0.04)],            2, 8, 64, 0.1, 250),
]


@pytest.mark.parametrize("name,specs,NT,MT,K,rate,M", CASES, ids=[c[0] for c in CASES])
def test_teacher_forced_parity(name, specs, NT, MT, K, rate, M):
    from ctrm_b200 import RoadmapBuilder

    w = O.OWeights.load(WEIGHTS)
    w.map_size = M
    instances = [_instance(*s) for s in specs]
    oins = [_oins(i) for i in instances]
    seed = 1000 + len(name)
    p = O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=rate, max_attempt=3, K=K)
    recs_all, trms_all = [], []
    for b, oi in enumerate(oins):
        recs = []
        trms_all.append(O.build_timed_roadmaps(oi, w, p, PhiloxTables(seed, 3 + b), recorder=recs))
        recs_all.append(recs)
    teacher = np.concatenate([teacher_table(r, oi.num_agents, K, NT, MT, key="SAMPLES_nn") for r, oi in zip(recs_all, oins)], axis=2)
    builder = RoadmapBuilder(WEIGHTS, 0, map_size=M)
    try:
        builder.upload(instances)
        # batch path with one warp per agent / one thread per agent, and the per-instance cluster kernel (within its limits):
        # the same bits for the roadmaps, the CVAE samples within 1e-5 of the oracle's
        inst_ok = K <= 8 and max(oi.num_agents for oi in oins) <= 64
        for mode in ((1, 2, 3) if inst_ok else (1, 2)):
            builder.set_step_kernel(mode)
            csr = builder.run(N_traj=NT, max_T=MT, max_attempt=3, merge_distance_rate=rate, K=K, seed=seed, instance_id_base=3,
                              tables=dict(teacher=teacher), trace=True)
            samples, _, loc_next = builder.read_trace()
            a0, worst = 0, 0.0
            for b, oi in enumerate(oins):
                assert_same_roadmaps(csr_to_arrays(csr, b), otrms_to_arrays(trms_all[b]))
                assert int(csr.cnt_collide[b]) == oi.cnt_continuous_collide
                N = oi.num_agents
                for r in recs_all[b]:
                    k, t = r["k_traj"], r["t"] - 1
                    assert np.array_equal(loc_next[k, t, a0:a0 + N], np.array(r["loc_next"]))
                    worst = max(worst, rel_err(samples[k, t, a0:a0 + N], r["SAMPLES_nn"].reshape(K, N, 3).transpose(1, 0, 2)))
                a0 += N
            assert worst <= 1e-5, worst
    finally:
        builder.close()


def test_randomize_indicator_parity():
    """randomize_indicator=True (learned_sampler.py:93-97): when the step's draw exceeds p_rw every row k*N + i takes a
    torch.randint indicator instead of the argmax.  Oracle with the same keyed draws: CVAE samples within 1e-5 at every
    executed step (they depend on the per-row indicators), roadmaps bit-identical under teacher forcing"""
    from ctrm_b200 import RoadmapBuilder

    w = O.OWeights.load(WEIGHTS)
    instances = [_instance(21, 9, 10, 6), _instance(22, 5, 6, 3)]
    oins = [_oins(i) for i in instances]
    NT, MT, K, rate, seed = 5, 24, 3, 0.1, 4242
    p = O.OParams(N_traj=NT, max_T=MT, merge_distance_rate=rate, max_attempt=3, K=K, randomize_indicator=True)
    recs_all, trms_all = [], []
    for b, oi in enumerate(oins):
        recs = []
        trms_all.append(O.build_timed_roadmaps(oi, w, p, PhiloxTables(seed, 11 + b), recorder=recs))
        recs_all.append(recs)
    n_random = sum(1 for recs in recs_all for r in recs if r["ind_logits"] is None)
    assert 0 < n_random < sum(len(r) for r in recs_all)  # both branches of learned_sampler.py:93 are exercised
    teacher = np.concatenate([teacher_table(r, oi.num_agents, K, NT, MT, key="SAMPLES_nn") for r, oi in zip(recs_all, oins)], axis=2)
    builder = RoadmapBuilder(WEIGHTS, 0)
    try:
        builder.upload(instances)
        csr = builder.run(N_traj=NT, max_T=MT, max_attempt=3, merge_distance_rate=rate, K=K, seed=seed, instance_id_base=11,
                          randomize_indicator=True, tables=dict(teacher=teacher), trace=True)
        samples, _, _ = builder.read_trace()
        a0, worst = 0, 0.0
        for b, oi in enumerate(oins):
            assert_same_roadmaps(csr_to_arrays(csr, b), otrms_to_arrays(trms_all[b]))
            N = oi.num_agents
            for r in recs_all[b]:
                k, t = r["k_traj"], r["t"] - 1
                worst = max(worst, rel_err(samples[k, t, a0:a0 + N], r["SAMPLES_nn"].reshape(K, N, 3).transpose(1, 0, 2)))
            a0 += N
        assert worst <= 1e-5, worst
        # free-running (no teacher, learned rows only) with the option on: runs and is repeatable
        c1 = builder.run(N_traj=NT, max_T=MT, merge_distance_rate=rate, seed=seed, randomize_indicator=True).copy()
        c2 = builder.run(N_traj=NT, max_T=MT, merge_distance_rate=rate, seed=seed, randomize_indicator=True).copy()
        assert np.array_equal(c1.pos.view(np.uint64), c2.pos.view(np.uint64)) and np.array_equal(c1.eidx, c2.eidx)
    finally:
        builder.close()


def test_api_error_paths():
    """bad arguments come back as negative status codes with a message, never as exceptions from native code"""
    import ctypes as C

    from ctrm_b200 import RoadmapBuilder, _lib

    lib = _lib.load()
    assert lib.ctrm_collide_segments(None, None, None, None, 5, None, None, None, None, None, None) == -1
    assert b"NULL" in lib.ctrm_last_error(None)
    assert lib.ctrm_fov_raster(None, None, 0, None, None, 0, 160, 19, None, None) == -1
    b = RoadmapBuilder(WEIGHTS, 0)
    try:
        with pytest.raises(_lib.CtrmError):
            b.run()  # nothing uploaded
        b.upload([_instance(11, 4, 5, 2)])
        with pytest.raises(_lib.CtrmError, match="N_traj"):
            b.run(N_traj=500)
        with pytest.raises(_lib.CtrmError, match="randomize_indicator"):  # recorded-stream tables hold no torch.randint draws
            b.run(N_traj=1, max_T=4, randomize_indicator=True,
                  tables=dict(gate=np.full((1, 4, 4), 0.5), rw=np.zeros((1, 4, 4, 3, 3))))
        nv = C.c_int64()
        assert lib.ctrm_result_sizes(b.ctx, C.byref(nv), None, None) == -3  # CTRM_ERR_STATE: no result yet
        csr = b.run(N_traj=0, max_T=8)  # zero roll-outs: only the start layer, extended once, plus goals
        assert csr.n_vertices > 0 and all(int(n) == 2 for n in csr.n_layers)
    finally:
        b.close()
